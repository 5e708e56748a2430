#!/bin/bash
# Round-2 GPU pass K: parity suite (hang-safe), bench c3 / c2 after the cp.async tail change, racecheck on the C3 modulator.
TAG=${1:-r02k}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 25 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== bench c3" ; timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e --no-latency 2>&1 | tail -n 1 | tee $OUT/${TAG}_bench_c3.json | cut -c1-200
echo "== bench c2" ; timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -n 1 | tee $OUT/${TAG}_bench_c2.json | cut -c1-200
timeout 240 compute-sanitizer --tool racecheck --print-limit 10 --log-file $OUT/${TAG}_sanitizer_racecheck_k1024.log python tools/sanitize_target.py 1024 > $OUT/${TAG}_sanitizer_racecheck_k1024_stdout.txt 2>&1
echo "rc=$?" >> $OUT/${TAG}_sanitizer_racecheck_k1024_stdout.txt; tail -n 2 $OUT/${TAG}_sanitizer_racecheck_k1024.log
