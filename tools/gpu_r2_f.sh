#!/bin/bash
# Round-2 GPU pass F: parity suite (hang-safe), chain rows.
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== chain rows" ; timeout 300 python tools/chain_bench.py rx > $OUT/${TAG}_chain_rx.jsonl 2> $OUT/${TAG}_chain_rx.err; tail -n 3 $OUT/${TAG}_chain_rx.err
timeout 240 compute-sanitizer --tool racecheck --print-limit 10 --log-file $OUT/${TAG}_sanitizer_racecheck_k128.log python tools/sanitize_target.py 128 > $OUT/${TAG}_sanitizer_racecheck_k128_stdout.txt 2>&1
echo "rc=$?" >> $OUT/${TAG}_sanitizer_racecheck_k128_stdout.txt; tail -n 2 $OUT/${TAG}_sanitizer_racecheck_k128.log
timeout 240 compute-sanitizer --tool racecheck --print-limit 10 --log-file $OUT/${TAG}_sanitizer_racecheck_k1024.log python tools/sanitize_target.py 1024 > $OUT/${TAG}_sanitizer_racecheck_k1024_stdout.txt 2>&1
echo "rc=$?" >> $OUT/${TAG}_sanitizer_racecheck_k1024_stdout.txt; tail -n 2 $OUT/${TAG}_sanitizer_racecheck_k1024.log
