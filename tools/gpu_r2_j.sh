#!/bin/bash
# Round-2 GPU pass J: parity suite (hang-safe), bench c3 + receiver rows after the modulator tail / templated equaliser changes.
TAG=${1:-r02j}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 25 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== bench c3" ; timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-latency 2>&1 | tail -n 1 | tee $OUT/${TAG}_bench_c3.json | cut -c1-200
echo "== chain rows" ; timeout 300 python tools/chain_bench.py rx > $OUT/${TAG}_chain_rx.jsonl 2> $OUT/${TAG}_chain_rx.err; tail -n 3 $OUT/${TAG}_chain_rx.err
