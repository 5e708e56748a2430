#!/usr/bin/env python3
"""Device-resident throughput of the chain entry points around the modulator/receiver
against their algorithmic bytes.  One JSON line per measurement; CUDA-event timing on the
handle's stream, buffers larger than L2.  usage (GPU box): python tools/chain_bench.py [tx] [rx] [est] [adv]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'gr-gfdm_b200'))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from gfdm_b200 import capi, design  # noqa: E402

lib = capi.load()
lib.set_device(0)
PEAK = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))).get('hbm_gbs', 6650.0)
stream = torch.cuda.Stream()


def timed(fn, steps=10, warmup=3):
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def report(what, shape, frames, ms, alg_bytes_frame, kernel, extra=None):
    gbs = alg_bytes_frame * frames / (ms * 1e-3) / 1e9
    line = {'what': what, 'shape': shape, 'frames': frames, 'ms': ms, 'frames_per_s': frames / (ms * 1e-3),
            'algorithmic_bytes_per_frame': alg_bytes_frame, 'achieved_gbs': gbs, 'peak_gbs': PEAK, 'frac': gbs / PEAK,
            'kernel': kernel}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def crand(rng, *shape):
    return (rng.standard_normal(shape, dtype=np.float32) + 1j * rng.standard_normal(shape, dtype=np.float32)).astype(np.complex64)


def tx_chain(M, K, A, cp, cs, frames, shifts=(0,)):
    """transmitter_kernel::generic_work per antenna: 8*(A*M + P+cp+N+cs) bytes per frame."""
    rng = np.random.default_rng(7)
    ramp = cs
    N = M * K
    taps = design.get_frequency_domain_filter('rrc', 0.5, M, K, 2)
    window = design.get_raised_cosine_ramp(ramp, design.get_window_len(cp, M, K, cs))
    smap = design.get_subcarrier_map(K, A, dc_free=True)
    P = 2 * K + cp + ramp
    pre = [crand(rng, P) for _ in shifts]
    tx = capi.Transmitter(M, K, A, cp, cs, ramp, smap, True, 2, taps, window, list(shifts), pre, lib=lib)
    tx.set_stream(stream.cuda_stream)
    pts = design.qam16_points().astype(np.complex64)
    d_in = torch.from_numpy(pts[rng.integers(0, 16, (frames, A * M))]).cuda()
    os_ = tx.output_vector_size()
    d_out = torch.empty((len(shifts), frames, os_), dtype=torch.complex64, device='cuda')
    alg = 8 * (A * M + len(shifts) * os_)
    for fused in (True, False):
        tx.set_chain_fusion(fused)
        ms = timed(lambda: tx.work_ptr(d_out.data_ptr(), d_in.data_ptr(), A * M, frames, all_antennas=True))
        report('transmitter_kernel chain (%s)' % ('one kernel' if fused else 'separate kernels'),
               'K=%d M=%d A=%d cp=%d cs=%d antennas=%d' % (K, M, A, cp, cs, len(shifts)), frames, ms, alg,
               tx.last_kernel(), {'msamples_per_s': frames * os_ * len(shifts) / (ms * 1e-3) / 1e6})
    tx.set_chain_fusion(True)


if __name__ == '__main__':
    what = sys.argv[1:] or ['tx']
    if 'tx' in what:
        tx_chain(9, 64, 52, 16, 8, 1 << 16)            # BASELINE configs[1]
        tx_chain(15, 1024, 832, 64, 32, 4096)           # headline shape with CP
        tx_chain(15, 256, 208, 32, 16, 1 << 14, (0, 16))
