#!/usr/bin/env python3
"""One small invocation of every hand-rolled fused kernel (mbarrier phases, cp.async.bulk loads / store groups, tensor
memory alloc / dealloc) -- the target of tools/gpu_sanitize.sh (compute-sanitizer memcheck / racecheck / synccheck)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'gr-gfdm_b200'))
os.environ.setdefault('HOME', '/tmp')
from gfdm_b200 import capi, design  # noqa: E402

lib = capi.load()
rng = np.random.default_rng(1)
only = sys.argv[1:]


def crand(*shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)


seen = []
for M, K, frames in ((5, 16, 300), (9, 64, 70), (15, 256, 9), (15, 1024, 3), (15, 2048, 3), (21, 128, 5)):
    if only and str(K) not in only:
        continue
    N, L = M * K, 2
    taps = design.get_frequency_domain_filter('rrc', .5, M, K, L)
    A = design.default_active_subcarriers(K)
    smap = design.get_subcarrier_map(K, A, dc_free=True)
    mod = capi.Modulator(M, K, L, taps, lib=lib)
    dem = capi.Demodulator(M, K, L, np.conj(taps), lib=lib)
    d = crand(frames, N)
    x = mod.modulate_batch(d); seen.append(mod.last_kernel())
    y = dem.demodulate_batch(x); seen.append(dem.last_kernel())
    eq = (1 + 0.1 * crand(frames, N)).astype(np.complex64)
    dem.demodulate_batch(x, eq); seen.append(dem.last_kernel() + '+eq')
    dem.fft_filter_downsample_batch(x) if hasattr(dem, 'fft_filter_downsample_batch') else None
    sm = capi.Symbol_mapper((design.qam16_points(), capi.DECISION_NEAREST), lib=lib)
    ch = rng.integers(0, 16, (frames, N)).astype(np.uint8)
    mod.modulate_chunks_batch(sm, ch); seen.append(mod.last_kernel())
    dem.demodulate_decide_batch(sm, x); seen.append(dem.last_kernel())
    for pts, rule in (capi.qpsk_constellation(), (design.qam16_points().astype(np.complex64), capi.DECISION_NEAREST)):
        adv = capi.Advanced_receiver(M, K, L, np.conj(taps), smap, 2, (pts, rule), 1, lib=lib)
        adv.demodulate_batch(x, eq); seen.append(adv.last_kernel())
    cp, cs = K // 4, K // 8
    window = design.get_raised_cosine_ramp(cs, design.get_window_len(cp, M, K, cs))
    pre = [crand(2 * K + cp + cs) for _ in range(2)]
    tx = capi.Transmitter(M, K, A, cp, cs, cs, smap, True, L, taps, window, [0, 2], pre, lib=lib)
    s = crand(frames, tx.input_vector_size())
    tx.work_all_batch(s); seen.append(tx.last_kernel())
    sh = capi.Burst_shaper(3, 5, 0.5 + 0.1j, lib=lib)
    tx.work_shaped_batch(sh, s); seen.append(tx.last_kernel() + '+shaper')
    est = capi.Preamble_channel_estimator(M, K, A, True, 1, crand(2 * K), lib=lib)
    est.estimate_frame_batch(crand(frames, 2 * K)); seen.append(est.last_kernel())
print('kernels exercised:')
for k in sorted(set(seen)):
    print('  ', k)
print('sanitize target done')
