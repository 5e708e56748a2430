#!/bin/bash
# session pass: parity, C3 bench, stage cycles, transmitter chain bench
TAG=${1:-s6a}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_gpu.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | tee $OUT/${TAG}_bench_c3.json
timeout 300 python tools/chain_bench.py tx 2>&1 | tail -8 | tee $OUT/${TAG}_chain.json
timeout 300 python tools/stage_profile.py c3 2>&1 | tail -30 | tee $OUT/${TAG}_stages.txt
