#!/bin/bash
# Round-2 GPU pass U: parked receiver with the steps reordered (0, 2, 1, 3), optional start offset of every second CTA.
TAG=${1:-r02u}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest K=2048" ; timeout 300 python -m pytest tests -m gpu -x -q -k "2048 or twopass" 2>&1 | tail -n 6 | tee $OUT/${TAG}_pytest_gpu.txt
for tool in memcheck racecheck; do
  timeout 200 compute-sanitizer --tool $tool --print-limit 10 --log-file $OUT/${TAG}_sanitizer_${tool}_k2048.log \
      python tools/sanitize_target.py 2048 > $OUT/${TAG}_sanitizer_${tool}_k2048_stdout.txt 2>&1
  tail -n 2 $OUT/${TAG}_sanitizer_${tool}_k2048.log; tail -n 1 $OUT/${TAG}_sanitizer_${tool}_k2048_stdout.txt
done
for st in 0 10000 20000 30000; do
  GFDM_TWOPASS_STAGGER=$st timeout 300 python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu --no-e2e --no-latency 2>&1 | tail -n 1 > $OUT/${TAG}_bench_c5_stagger$st.json
done
python - <<'PY'
import json
for st in (0, 10000, 20000, 30000):
    d = json.loads(open('gpurun_out/r02u_bench_c5_stagger%d.json' % st).read().strip().splitlines()[-1])
    print('stagger', st, d['roofline']['kernel_ms'], round(d['roofline']['chain_frac'], 4))
PY
timeout 200 python tools/stage_profile.py c5 2048 2>&1 | tail -n 12 | tee $OUT/${TAG}_stage_cycles_c5_parked.txt
