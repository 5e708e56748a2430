#!/bin/bash
# Multi-GPU pass (gpurun --gpus N): link ceiling with n = 1..N GPUs copying at once (threads of one process and separate
# processes), bench.py at N ranks launched exactly like the driver does, the C5 sweep, the single-process multi_gpu driver.
# Usage: bash tools/gpu_multi.sh <N> [tag]
N=${1:-2}
TAG=${2:-r02_n$N}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
nproc > $OUT/${TAG}_nproc.txt; numactl -H >> $OUT/${TAG}_nproc.txt 2>&1; lscpu | head -25 >> $OUT/${TAG}_nproc.txt 2>&1
echo "== link ceiling" ; timeout 300 tools/pcie_ceiling --gpus $N --secs 0.6 2>&1 | tee $OUT/${TAG}_pcie_ceiling.jsonl
PORT=29511
for n in $(seq 1 $N); do
  if [ $n -eq 1 ] || [ $n -eq 2 ] || [ $n -eq 4 ] || [ $n -eq 8 ]; then
    echo "== bench c3 N=$n"
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu 2>&1 | tail -1 | tee $OUT/${TAG}_bench_c3_n$n.json
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT \
          bench.py --gpus $n --steps 20 --warmup 3 --no-cpu 2>&1 | grep '^{' | tail -1 | tee $OUT/${TAG}_bench_c3_n$n.json
      PORT=$((PORT+1))
    fi
  fi
done
echo "== bench c5 sweep N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --workload c5 --sweep --gpus $N --steps 5 --no-cpu 2>&1 | grep '^{' | tail -1 | tee $OUT/${TAG}_bench_c5_sweep_n$N.json
echo "== reference arm N=$N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((PORT+1)) \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | grep '^{' | tail -1 | tee $OUT/${TAG}_bench_ref_n$N.json
echo "== single-process multi_gpu driver"
g++ -std=c++17 -O1 -Iinclude tests/cpp/multi_gpu_probe.cc -o /tmp/mgp -Lgr-gfdm_b200/lib -lgfdm_b200 -Wl,-rpath,$PWD/gr-gfdm_b200/lib -lpthread \
  && for w in 1 $N; do timeout 120 /tmp/mgp $w 4096; done 2>&1 | tee $OUT/${TAG}_multi_gpu_probe.txt
ls -la $OUT | tail -12
