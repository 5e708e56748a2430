#!/bin/bash
# Round-2 GPU pass L: parity suite (hang-safe) incl. the generic shared-memory kernels, per-shape rows, racecheck on them.
TAG=${1:-r02m}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 25 | tee $OUT/${TAG}_pytest_gpu.txt
echo "== shape rows" ; timeout 400 python tools/chain_bench.py shapes > $OUT/${TAG}_chain_shapes.jsonl 2> $OUT/${TAG}_chain_shapes.err; tail -n 3 $OUT/${TAG}_chain_shapes.err
timeout 200 compute-sanitizer --tool racecheck --print-limit 10 --log-file $OUT/${TAG}_sanitizer_racecheck_generic.log python - > $OUT/${TAG}_sanitizer_racecheck_generic_stdout.txt 2>&1 <<'PY'
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
tail -n 2 $OUT/${TAG}_sanitizer_racecheck_generic.log; tail -n 4 $OUT/${TAG}_sanitizer_racecheck_generic_stdout.txt
