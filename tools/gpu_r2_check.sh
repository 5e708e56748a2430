#!/bin/bash
# Round-2 GPU pass: parity tests, smoke, bench lines (all named workloads), link ceiling.  Usage: bash tools/gpu_r2_check.sh [tag]
TAG=${1:-r02a}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.txt
echo "== bench c3" ; timeout 900 python bench.py --steps 20 --warmup 3 2>&1 | tail -2 | tee $OUT/${TAG}_bench_c3.json
for wl in c2 c4 c5 c1; do
  echo "== bench $wl" ; timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --cpu-seconds 6 2>&1 | tail -1 | tee $OUT/${TAG}_bench_$wl.json
done
echo "== bench c5 sweep" ; timeout 900 python bench.py --workload c5 --sweep --steps 5 --no-cpu 2>&1 | tail -1 | tee $OUT/${TAG}_bench_c5_sweep.json
echo "== link ceiling" ; timeout 300 tools/pcie_ceiling --gpus 1 --secs 0.5 2>&1 | tee $OUT/${TAG}_pcie_ceiling.jsonl
ls -la $OUT | tail -20
