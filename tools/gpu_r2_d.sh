#!/bin/bash
# Round-2 GPU pass D: parity suite (hang-safe), bench c4/c3, chain rows, C5 stage profile, sanitizer for K=128.
TAG=${1:-r02d}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== bench c4" ; timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $OUT/${TAG}_bench_c4.json
echo "== bench c3" ; timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $OUT/${TAG}_bench_c3.json
echo "== chain rows" ; timeout 300 python tools/chain_bench.py rx > $OUT/${TAG}_chain_rx.jsonl 2> $OUT/${TAG}_chain_rx.err; tail -3 $OUT/${TAG}_chain_rx.err
echo "== stage profile c5" ; timeout 200 python tools/stage_profile.py c5 2048 2>&1 | tee $OUT/${TAG}_stage_cycles_c5.txt
echo "== stage profile c3" ; timeout 200 python tools/stage_profile.py c3 4096 2>&1 | tee $OUT/${TAG}_stage_cycles_c3.txt
for tool in memcheck racecheck; do
  timeout 240 compute-sanitizer --tool $tool --print-limit 10 --log-file $OUT/${TAG}_sanitizer_${tool}_k128.log python tools/sanitize_target.py 128 > $OUT/${TAG}_sanitizer_${tool}_k128_stdout.txt 2>&1
  echo "rc=$?" >> $OUT/${TAG}_sanitizer_${tool}_k128_stdout.txt; tail -2 $OUT/${TAG}_sanitizer_${tool}_k128.log
done
