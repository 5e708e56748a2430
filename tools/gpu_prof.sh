#!/bin/bash
# parity (no -x) + ncu full capture of the C3 fused kernels
TAG=${1:-p}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee $OUT/${TAG}_pytest_gpu.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_ -s 6 -c 2 -f -o $OUT/${TAG}_fused \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > $OUT/${TAG}_full_bench.log 2>&1
ls -la $OUT | tail -5
