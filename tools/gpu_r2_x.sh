#!/bin/bash
# Round-2 GPU pass X: compute-sanitizer on the final build for the shapes whose kernels changed last (transmitter chain store loop:
# K=64 and K=1024 families; generic kernels), hang-safe.
TAG=${1:-r02x}
OUT=gpurun_out
mkdir -p $OUT
for tool in memcheck racecheck; do
  for k in 64 1024; do
    timeout 200 compute-sanitizer --tool $tool --print-limit 10 --log-file $OUT/${TAG}_sanitizer_${tool}_k$k.log \
        python tools/sanitize_target.py $k > $OUT/${TAG}_sanitizer_${tool}_k${k}_stdout.txt 2>&1
    echo "$tool K=$k: $(tail -n 1 $OUT/${TAG}_sanitizer_${tool}_k$k.log)"
  done
done
timeout 200 compute-sanitizer --tool memcheck --print-limit 10 --log-file $OUT/${TAG}_sanitizer_memcheck_generic.log python - > $OUT/${TAG}_sanitizer_memcheck_generic_stdout.txt 2>&1 <<'PY'
import sys, os
sys.path.insert(0, 'gr-gfdm_b200'); os.environ.setdefault('HOME', '/tmp')
import numpy as np
from gfdm_b200 import capi, design
lib = capi.load(); rng = np.random.default_rng(1)
for M, K, L in ((25, 96, 2), (127, 16, 4), (6, 12, 3), (16, 96, 2), (11, 416, 2), (13, 840, 2)):
    taps = design.get_frequency_domain_filter('rrc', .5, M, K, L)
    d = (rng.standard_normal((5, M * K)) + 1j * rng.standard_normal((5, M * K))).astype(np.complex64)
    mod, dem = capi.Modulator(M, K, L, taps, lib=lib), capi.Demodulator(M, K, L, np.conj(taps), lib=lib)
    x = mod.modulate_batch(d); y = dem.demodulate_batch(x, (1 + 0 * d).astype(np.complex64)); r = dem.fft_filter_downsample_batch(x)
    print(M, K, mod.last_kernel(), dem.last_kernel())
PY
echo "memcheck generic: $(tail -n 1 $OUT/${TAG}_sanitizer_memcheck_generic.log)"; tail -n 3 $OUT/${TAG}_sanitizer_memcheck_generic_stdout.txt
