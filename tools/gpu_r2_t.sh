#!/bin/bash
# Round-2 GPU pass T: two-pass receiver with tensor-memory parking -- parity at K=2048 (hang-safe), sanitizers, A/B, stage profile.
TAG=${1:-r02t}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest K=2048" ; timeout 300 python -m pytest tests -m gpu -x -q -k "2048 or twopass" 2>&1 | tail -n 6 | tee $OUT/${TAG}_pytest_gpu.txt
for tool in memcheck racecheck; do
  timeout 200 compute-sanitizer --tool $tool --print-limit 10 --log-file $OUT/${TAG}_sanitizer_${tool}_k2048.log \
      python tools/sanitize_target.py 2048 > $OUT/${TAG}_sanitizer_${tool}_k2048_stdout.txt 2>&1
  tail -n 2 $OUT/${TAG}_sanitizer_${tool}_k2048.log; tail -n 2 $OUT/${TAG}_sanitizer_${tool}_k2048_stdout.txt
done
echo "== bench c5 parked" ; timeout 300 python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu --no-e2e --no-latency 2>&1 | tail -n 1 | tee $OUT/${TAG}_bench_c5_parked.json | cut -c1-100
echo "== bench c5 first versions" ; GFDM_RX2_REREAD=1 GFDM_MOD2_SCRATCH=1 timeout 300 python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu --no-e2e --no-latency 2>&1 | tail -n 1 | tee $OUT/${TAG}_bench_c5_first.json | cut -c1-100
python - <<'PY'
import json
for v in ('parked', 'first'):
    d = json.loads(open('gpurun_out/r02t_bench_c5_%s.json' % v).read().strip().splitlines()[-1])
    print(v, d['roofline']['kernel_ms'], d['roofline']['chain_frac'], d['config'].get('kernels'))
PY
timeout 200 python tools/stage_profile.py c5 2048 2>&1 | tee $OUT/${TAG}_stage_cycles_c5_parked.txt
