#!/bin/bash
# One GPU-box evidence pass: parity tests, smoke, bench lines of every BASELINE config (own + reference arm), per-kernel rows,
# ncu launch list, one ncu full capture of the two headline kernels.  Usage (under gpurun): bash tools/gpu_check.sh [tag]
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 15 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3 | tee $OUT/${TAG}_smoke.txt
echo "== bench c3" ; timeout 600 python bench.py --steps 20 --warmup 3 2>&1 | tail -n 1 | tee $OUT/${TAG}_bench_c3.json | cut -c1-300
echo "== bench reference" ; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -n 1 | tee $OUT/${TAG}_bench_ref.json | cut -c1-300
for wl in c1 c2 c4 c5; do
  echo "== bench $wl" ; timeout 400 python bench.py --workload $wl --steps 10 --warmup 3 --cpu-seconds 6 2>&1 | tail -n 1 | tee $OUT/${TAG}_bench_$wl.json | cut -c1-200
done
echo "== bench c5 sweep" ; timeout 600 python bench.py --workload c5 --sweep --steps 5 --no-cpu 2>&1 | tail -n 1 | tee $OUT/${TAG}_bench_c5_sweep_n1.json | cut -c1-200
echo "== per-row chain bench"
timeout 600 python tools/chain_bench.py tx rx copy next shapes > $OUT/${TAG}_chain_bench.jsonl 2> $OUT/${TAG}_chain_bench.err; tail -n 2 $OUT/${TAG}_chain_bench.err
echo "== ncu launch list"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-latency > $OUT/${TAG}_launches_bench.log 2>&1
echo "== ncu full"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_ -s 6 -c 2 -f -o $OUT/${TAG}_fused \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-latency > $OUT/${TAG}_full_bench.log 2>&1
rm -f $OUT/${TAG}_fused.ncu-rep.tmp
ncu -i $OUT/${TAG}_fused.ncu-rep --page raw --csv > $OUT/${TAG}_fused_raw.csv 2>/dev/null && python tools/ncu_summary.py $OUT/${TAG}_fused_raw.csv > $OUT/${TAG}_c3_fused_ncu_full.txt
timeout 200 python tools/stage_profile.py c5 2048 > $OUT/${TAG}_stage_cycles_c5.txt 2>&1
ls -la $OUT | tail -n 20
