#!/bin/bash
# Round-2 GPU pass B: full parity suite (incl. block shims, shaper, legacy 2-D, multi-GPU driver), sanitizers, bench c3.
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/${TAG}_pytest_gpu.txt
bash tools/gpu_sanitize.sh $TAG
echo "== bench c3" ; timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -2 | tee $OUT/${TAG}_bench_c3.json
