#!/bin/bash
# Round-2 GPU pass G: parity suite (hang-safe), bench c5 / c3 after the receiver-table and two-pass output changes.
TAG=${1:-r02g}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 25 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== bench c5" ; timeout 300 python bench.py --workload c5 --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -n 1 | tee $OUT/${TAG}_bench_c5.json | cut -c1-300
echo "== bench c3" ; timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -n 1 | tee $OUT/${TAG}_bench_c3.json | cut -c1-300
