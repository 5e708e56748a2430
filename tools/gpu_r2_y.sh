#!/bin/bash
# Round-2 GPU pass Y (gpurun --gpus 2): the C5 sweep and the C5 bench line on 2 ranks with the tensor-memory two-pass kernels,
# launched like the driver does.
OUT=gpurun_out
mkdir -p $OUT
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --workload c5 --sweep --gpus 2 --steps 5 --no-cpu 2>&1 | grep '^{' | tail -n 1 | tee $OUT/r02y_bench_c5_sweep_n2.json | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 \
    bench.py --workload c5 --gpus 2 --steps 20 --warmup 3 --no-cpu 2>&1 | grep '^{' | tail -n 1 | tee $OUT/r02y_bench_c5_n2.json | cut -c1-200
