#!/bin/bash
# Round-2 GPU pass V: equalising variant of the parked two-pass receiver -- parity at K=2048 (hang-safe), sanitizers, row timing.
TAG=${1:-r02v}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest K=2048" ; timeout 400 python -m pytest tests -m gpu -x -q -k "2048 or twopass" 2>&1 | tail -n 12 | tee $OUT/${TAG}_pytest_gpu.txt
for tool in memcheck racecheck; do
  timeout 300 compute-sanitizer --tool $tool --print-limit 10 --log-file $OUT/${TAG}_sanitizer_${tool}_k2048.log \
      python tools/sanitize_target.py 2048 > $OUT/${TAG}_sanitizer_${tool}_k2048_stdout.txt 2>&1
  tail -n 2 $OUT/${TAG}_sanitizer_${tool}_k2048.log; tail -n 3 $OUT/${TAG}_sanitizer_${tool}_k2048_stdout.txt
done
timeout 300 python tools/chain_bench.py rx2048 2>&1 | tee $OUT/${TAG}_chain_rx2048.jsonl | cut -c1-330
timeout 300 python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu --no-e2e --no-latency 2>&1 | tail -n 1 | tee $OUT/${TAG}_bench_c5.json | cut -c1-120
