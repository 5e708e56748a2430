#!/bin/bash
# quick GPU pass while iterating on the fused kernels: parity + bench C1..C4 + stage cycles
TAG=${1:-q}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_gpu.txt
for wl in c3 c4 c2 c1; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | tee $OUT/${TAG}_bench_$wl.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('$wl', '%.3g frames/s' % d['value'], 'mod %.4f ms rx %.4f ms' % (r['kernel_ms']['modulator'], r['kernel_ms']['receiver']), 'chain frac %.3f' % r['chain_frac'], d['config']['kernels'])"
done
[ -f gr-gfdm_b200/lib/libgfdm_b200_prof.so ] && timeout 300 python tools/stage_profile.py c3 2>&1 | tail -25 | tee $OUT/${TAG}_stages.txt
