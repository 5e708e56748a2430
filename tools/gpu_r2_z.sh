#!/bin/bash
# Round-2 GPU pass Z: A/B of a receiver output variant (one bulk store per group instead of one per item): bench lines, parity suite.
TAG=${1:-r02z}
OUT=gpurun_out
mkdir -p $OUT
for wl in c3 c1 c2 c4; do
  timeout 300 python bench.py --workload $wl --steps 20 --warmup 3 --no-cpu --no-e2e --no-latency 2>&1 | tail -n 1 > $OUT/${TAG}_bench_$wl.json
  python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_$wl.json').read().strip().splitlines()[-1]); print('$wl', d['roofline']['kernel_ms'], round(d['roofline']['chain_frac'],4), d['clocks']['sm_mhz'])"
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3 | tee $OUT/${TAG}_pytest_gpu.txt
