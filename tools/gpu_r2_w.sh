#!/bin/bash
# Round-2 GPU pass W: quick A/B of a scheduling change in the two-pass modulator (bench line + stage cycles), K=2048 parity.
TAG=${1:-r02w}
OUT=gpurun_out
mkdir -p $OUT
timeout 400 python -m pytest tests -m gpu -x -q -k "2048 or twopass" 2>&1 | tail -n 3 | tee $OUT/${TAG}_pytest_gpu.txt
timeout 300 python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu --no-e2e --no-latency 2>&1 | tail -n 1 > $OUT/${TAG}_bench_c5.json
python -c "
import json; d=json.loads(open('gpurun_out/${TAG}_bench_c5.json').read().strip().splitlines()[-1]); print(d['roofline']['kernel_ms'], d['roofline']['chain_frac'])"
timeout 200 python tools/stage_profile.py c5 2048 2>&1 | head -12 | tee $OUT/${TAG}_stage_cycles_c5.txt
