#!/bin/bash
# 8-GPU pass (gpurun --gpus 8): link ceiling for n = 1..8 GPUs copying at once, bench.py at 4 and 8 ranks launched like the
# driver does, the C5 sweep on 8 GPUs, the reference arm, the single-process multi_gpu driver on 8 devices.
TAG=${1:-r02_n8}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
(nproc; lscpu | head -25; free -g) > $OUT/${TAG}_host.txt 2>&1
echo "== link ceiling" ; timeout 300 tools/pcie_ceiling --gpus 8 --secs 0.5 2>&1 | tee $OUT/${TAG}_pcie_ceiling.jsonl | tail -n 6
PORT=29611
for n in 4 8; do
  echo "== bench c3 N=$n"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $PORT \
      bench.py --gpus $n --steps 20 --warmup 3 --no-cpu 2>&1 | grep '^{' | tail -n 1 | tee $OUT/${TAG}_bench_c3_n$n.json | cut -c1-200
  PORT=$((PORT+1))
done
echo "== bench c5 sweep N=8"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT \
    bench.py --workload c5 --sweep --gpus 8 --steps 5 --no-cpu 2>&1 | grep '^{' | tail -n 1 | tee $OUT/${TAG}_bench_c5_sweep_n8.json | cut -c1-200
echo "== reference arm N=8"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((PORT+1)) \
    bench.py --impl reference --gpus 8 --steps 2 --warmup 1 2>&1 | grep '^{' | tail -n 1 | tee $OUT/${TAG}_bench_ref_n8.json | cut -c1-200
echo "== single-process multi_gpu driver"
g++ -std=c++17 -O1 -Iinclude tests/cpp/multi_gpu_probe.cc -o /tmp/mgp -Lgr-gfdm_b200/lib -lgfdm_b200 -Wl,-rpath,$PWD/gr-gfdm_b200/lib -lpthread \
  && for w in 1 8; do timeout 120 /tmp/mgp $w 8192; done 2>&1 | tee $OUT/${TAG}_multi_gpu_probe.txt
