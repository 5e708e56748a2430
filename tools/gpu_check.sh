#!/bin/bash
# One GPU-box pass: parity tests, smoke, bench (own + reference arm), ncu launch list, ncu full capture.
# Usage (from the repo root, under gpurun): bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.txt
echo "== bench c3" ; timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -2 | tee $OUT/${TAG}_bench_c3.json
echo "== bench reference" ; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee $OUT/${TAG}_bench_ref.json
for wl in c1 c2 c4 c5; do
  echo "== bench $wl" ; timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | tee $OUT/${TAG}_bench_$wl.json
done
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/${TAG}_launches_bench.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_ -s 6 -c 2 -f -o $OUT/${TAG}_fused \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > $OUT/${TAG}_full_bench.log 2>&1
ls -la $OUT
echo "== per-row chain bench"
timeout 600 python tools/chain_bench.py tx rx copy next > $OUT/${TAG}_chain_bench.jsonl 2> $OUT/${TAG}_chain_bench.err
rm -f $OUT/${TAG}_fused.ncu-rep.tmp
# the full ncu report is summarised on the box as well (tools/ncu_summary.py), in case it is too large to travel
ncu -i $OUT/${TAG}_fused.ncu-rep --page raw --csv > $OUT/${TAG}_fused_raw.csv 2>/dev/null && python tools/ncu_summary.py $OUT/${TAG}_fused_raw.csv > $OUT/${TAG}_c3_fused_ncu_full.txt
ls -la $OUT | tail -30
