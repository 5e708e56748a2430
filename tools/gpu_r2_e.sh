#!/bin/bash
# Round-2 GPU pass E: parity suite (hang-safe), chain rows, bench c2/c4, sanitizer for the shapes with tensor-memory twiddles.
TAG=${1:-r02e}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== chain rows" ; timeout 300 python tools/chain_bench.py rx > $OUT/${TAG}_chain_rx.jsonl 2> $OUT/${TAG}_chain_rx.err; tail -3 $OUT/${TAG}_chain_rx.err
echo "== bench c2" ; timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | tee $OUT/${TAG}_bench_c2.json
echo "== bench c4" ; timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | tee $OUT/${TAG}_bench_c4.json
for tool in memcheck racecheck; do
  timeout 240 compute-sanitizer --tool $tool --print-limit 10 --log-file $OUT/${TAG}_sanitizer_${tool}_k128.log python tools/sanitize_target.py 128 > $OUT/${TAG}_sanitizer_${tool}_k128_stdout.txt 2>&1
  echo "rc=$?" >> $OUT/${TAG}_sanitizer_${tool}_k128_stdout.txt; tail -n 2 $OUT/${TAG}_sanitizer_${tool}_k128.log
done
