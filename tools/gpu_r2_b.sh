#!/bin/bash
# Round-2 GPU pass: full parity suite (incl. block shims, shaper, legacy 2-D, multi-GPU driver, open shape table), bench c3/c2,
# sanitizers last (bounded per invocation).
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -25 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== bench c3" ; timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -2 | tee $OUT/${TAG}_bench_c3.json
echo "== bench c2" ; timeout 600 python bench.py --workload c2 --steps 10 --warmup 3 --cpu-seconds 6 2>&1 | tail -1 | tee $OUT/${TAG}_bench_c2.json
bash tools/gpu_sanitize.sh $TAG
