#!/usr/bin/env python3
"""Device-resident throughput of the chain entry points around the modulator/receiver
against their algorithmic bytes.  One JSON line per measurement; CUDA-event timing on the
handle's stream, buffers larger than L2.  usage (GPU box): python tools/chain_bench.py [tx] [rx] [copy] [next]"""
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


STEPS = int(os.environ.get('CHAIN_STEPS', '10'))   # 1 under ncu
WARMUP = int(os.environ.get('CHAIN_WARMUP', '3'))


def timed(fn, steps=None, warmup=None):
    steps = STEPS if steps is None else steps
    warmup = WARMUP if warmup is None else warmup
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


def rx_eq_row(M, K, frames):
    """Equalising receiver alone (24N bytes/frame) -- the K = 2048 shape, whose frame is larger than shared memory."""
    rng = np.random.default_rng(12)
    N = M * K
    taps = design.get_frequency_domain_filter('rrc', 0.5, M, K, 2)
    d_x = torch.from_numpy(crand(rng, frames, N)).cuda()
    d_h = torch.from_numpy((1 + 0.3 * crand(rng, frames, N)).astype(np.complex64)).cuda()
    d_y = torch.empty_like(d_x)
    rx = capi.Demodulator(M, K, 2, np.conj(taps), lib=lib)
    rx.set_stream(stream.cuda_stream)
    ms = timed(lambda: rx.demodulate_ptr(d_y.data_ptr(), d_x.data_ptr(), d_h.data_ptr(), frames))
    report('receiver_kernel_cc::generic_work_equalize', 'K=%d M=%d' % (K, M), frames, ms, 24 * N, rx.last_kernel())


def rx_rows(M, K, A, frames, ic_iter=4):
    """Receiver-side rows at one shape: equalising receiver (24N bytes/frame), preamble channel estimator
    (8*(2K+N)), advanced receiver with `ic_iter` SIC iterations resident on the SM (24N regardless of ic_iter)."""
    rng = np.random.default_rng(11)
    N = M * K
    shape = 'K=%d M=%d A=%d' % (K, M, A)
    taps = design.get_frequency_domain_filter('rrc', 0.5, M, K, 2)
    smap = design.get_subcarrier_map(K, A, dc_free=True)
    d_x = torch.from_numpy(crand(rng, frames, N)).cuda()
    d_h = torch.from_numpy((1 + 0.3 * crand(rng, frames, N)).astype(np.complex64)).cuda()
    d_y = torch.empty_like(d_x)
    rx = capi.Demodulator(M, K, 2, np.conj(taps), lib=lib)
    rx.set_stream(stream.cuda_stream)
    ms = timed(lambda: rx.demodulate_ptr(d_y.data_ptr(), d_x.data_ptr(), d_h.data_ptr(), frames))
    report('receiver_kernel_cc::generic_work_equalize', shape, frames, ms, 24 * N, rx.last_kernel())
    for it in (0, ic_iter):
        adv = capi.Advanced_receiver(M, K, 2, np.conj(taps), smap, it, capi.qpsk_constellation(), 0, lib=lib)
        adv.set_stream(stream.cuda_stream)
        ms = timed(lambda: adv.demodulate_ptr(d_y.data_ptr(), d_x.data_ptr(), d_h.data_ptr(), frames))
        report('advanced_receiver_kernel_cc::generic_work_equalize, %d SIC iterations' % it, shape, frames, ms, 24 * N,
               adv.last_kernel())
    # 16-QAM with nearest-point decisions needs a realistic signal (the decision cost depends on where the soft symbols
    # fall): mapped 16-QAM symbols through the modulator, channel = the equaliser's input, a little noise
    pts16 = design.qam16_points().astype(np.complex64)
    mp = capi.Resource_mapper(M, K, A, smap, True, lib=lib)
    mod = capi.Modulator(M, K, 2, taps, lib=lib)
    d_s = torch.from_numpy(pts16[rng.integers(0, 16, (frames, M * A))]).cuda()
    d_g = torch.empty_like(d_x)
    d_x16 = torch.empty_like(d_x)
    mp.map_ptr(d_g.data_ptr(), d_s.data_ptr(), M * A, frames)
    mp.sync()
    mod.modulate_ptr(d_x16.data_ptr(), d_g.data_ptr(), frames)
    mod.sync()
    d_x16 = torch.fft.ifft(torch.fft.fft(d_x16, dim=1) * d_h, dim=1)
    d_x16 = (d_x16 + 0.01 * torch.randn_like(d_x16)).to(torch.complex64).contiguous()
    adv = capi.Advanced_receiver(M, K, 2, np.conj(taps), smap, ic_iter, (pts16, capi.DECISION_NEAREST), 0, lib=lib)
    adv.set_stream(stream.cuda_stream)
    ms = timed(lambda: adv.demodulate_ptr(d_y.data_ptr(), d_x16.data_ptr(), d_h.data_ptr(), frames))
    report('advanced_receiver_kernel_cc::generic_work_equalize, %d SIC iterations, 16-QAM nearest-point decisions' % ic_iter, shape,
           frames, ms, 24 * N, adv.last_kernel())
    preamble = crand(rng, 2 * K)
    est = capi.Preamble_channel_estimator(M, K, A, True, 0, preamble, lib=lib)
    est.set_stream(stream.cuda_stream)
    d_p = torch.from_numpy(crand(rng, frames, 2 * K)).cuda()
    ms = timed(lambda: est.estimate_frame_ptr(d_h.data_ptr(), d_p.data_ptr(), frames))
    report('preamble_channel_estimator_cc::estimate_frame', shape, frames, ms, 8 * (2 * K + N), est.last_kernel())


def copy_rows(M, K, A, cp, cs, frames):
    """Integer / copy rows: resource mapper, demapper, cyclic prefixer (add with window, remove), remove_prefix."""
    rng = np.random.default_rng(13)
    N = M * K
    shape = 'K=%d M=%d A=%d cp=%d cs=%d' % (K, M, A, cp, cs)
    smap = design.get_subcarrier_map(K, A, dc_free=True)
    mp = capi.Resource_mapper(M, K, A, smap, True, lib=lib)
    mp.set_stream(stream.cuda_stream)
    d_c = torch.from_numpy(crand(rng, frames, A * M)).cuda()
    d_g = torch.empty((frames, N), dtype=torch.complex64, device='cuda')
    ms = timed(lambda: mp.map_ptr(d_g.data_ptr(), d_c.data_ptr(), A * M, frames))
    report('resource_mapper_kernel_cc::map_to_resources', shape, frames, ms, 8 * (A * M + N), mp.last_kernel())
    ms = timed(lambda: mp.demap_ptr(d_c.data_ptr(), d_g.data_ptr(), A * M, frames))
    report('resource_mapper_kernel_cc::demap_from_resources', shape, frames, ms, 8 * (2 * A * M), mp.last_kernel())
    W = N + cp + cs
    window = design.get_raised_cosine_ramp(cs, W)
    pf = capi.Cyclic_prefixer(N, cp, cs, cs, window, lib=lib)
    pf.set_stream(stream.cuda_stream)
    d_f = torch.empty((frames, W), dtype=torch.complex64, device='cuda')
    ms = timed(lambda: pf.add_ptr(d_f.data_ptr(), d_g.data_ptr(), 0, frames))
    report('add_cyclic_prefix_cc::add_cyclic_prefix', shape, frames, ms, 8 * (N + W), pf.last_kernel())
    ms = timed(lambda: pf.remove_ptr(d_g.data_ptr(), d_f.data_ptr(), frames))
    report('add_cyclic_prefix_cc::remove_cyclic_prefix', shape, frames, ms, 8 * (2 * N), pf.last_kernel())
    rp = capi.Remove_prefix(W, N, cp, lib=lib)
    rp.set_stream(stream.cuda_stream)
    ms = timed(lambda: rp.work_ptr(d_g.data_ptr(), d_f.data_ptr(), frames))
    report('remove_prefix_cc', shape, frames, ms, 8 * (2 * N), rp.last_kernel())


def next_rows(M, K, frames):
    """Rows either side of the path: symbol mapping kernels, burst extraction, byte-wide chain entries."""
    rng = np.random.default_rng(17)
    N = M * K
    n = frames * N
    shape = 'K=%d M=%d' % (K, M)
    sm = capi.Symbol_mapper((design.qam16_points(), capi.DECISION_NEAREST), lib=lib)
    sm.set_stream(stream.cuda_stream)
    d_ch = torch.randint(0, 16, (frames, N), dtype=torch.uint8, device='cuda')
    d_s = torch.empty((frames, N), dtype=torch.complex64, device='cuda')
    ms = timed(lambda: sm.map_chunks_ptr(d_s.data_ptr(), d_ch.data_ptr(), n))
    report('symbol mapper: chunks -> points', shape, frames, ms, 9 * N, sm.last_kernel())
    d_s.add_(0.1 * torch.randn_like(d_s))
    ms = timed(lambda: sm.decide_ptr(d_ch.data_ptr(), d_s.data_ptr(), n))
    report('symbol mapper: hard decisions', shape, frames, ms, 9 * N, sm.last_kernel())
    taps = design.get_frequency_domain_filter('rrc', 0.5, M, K, 2)
    mod, rx = capi.Modulator(M, K, 2, taps, lib=lib), capi.Demodulator(M, K, 2, np.conj(taps), lib=lib)
    mod.set_stream(stream.cuda_stream)
    rx.set_stream(stream.cuda_stream)
    d_x = torch.empty_like(d_s)
    ms = timed(lambda: mod.modulate_chunks_ptr(sm, d_x.data_ptr(), d_ch.data_ptr(), frames))
    report('modulator_kernel_cc from chunks', shape, frames, ms, 9 * N, mod.last_kernel())
    ms = timed(lambda: rx.demodulate_decide_ptr(sm, d_ch.data_ptr(), d_x.data_ptr(), 0, frames))
    report('receiver_kernel_cc to hard decisions', shape, frames, ms, 9 * N, rx.last_kernel())
    # extract_burst: one burst per frame-sized slot of a long stream, backoff 8, with and without CFO correction
    burst = N + 96
    starts = np.arange(frames, dtype=np.int64) * (burst + 32) + 40
    stream_len = int(starts[-1] + burst + 8)
    d_in = torch.from_numpy(crand(rng, stream_len)).cuda()
    d_b = torch.empty((frames, burst), dtype=torch.complex64, device='cuda')
    scales = rng.uniform(0.5, 2, frames).astype(np.float32)
    rots = np.exp(1j * rng.uniform(-1e-3, 1e-3, frames)).astype(np.complex64)
    for cfo in (False, True):
        eb = capi.Extract_burst(burst, 8, cfo, lib=lib)
        eb.set_stream(stream.cuda_stream)
        ms = timed(lambda: eb.work_ptr(d_b.data_ptr(), frames, d_in.data_ptr(), stream_len, starts, scales, rots))
        report('extract_burst_cc (%s)' % ('scale + CFO rotation' if cfo else 'scale only'), 'burst_len=%d' % burst, frames, ms,
               16 * burst, eb.last_kernel())


def shape_rows():
    """modulate -> demodulate on shapes beyond the BASELINE ones: entries of the open fused-shape table (the reference's own
    test shapes M=21/K=128, M=9/K=32, ...) and shapes outside it, which run the frame-resident generic kernels."""
    rng = np.random.default_rng(19)
    for M, K, L in ((21, 128, 2), (9, 32, 2), (15, 128, 2), (9, 512, 2), (15, 512, 2), (5, 1024, 2), (21, 512, 2), (7, 256, 2),
                    (25, 96, 2), (127, 16, 4), (15, 96, 2), (16, 96, 2), (11, 416, 2), (13, 840, 2)):
        N = M * K
        frames = max(256, (1 << 26) // N)       # ~0.5 GB per buffer: larger than L2
        taps = design.get_frequency_domain_filter('rrc', 0.5, M, K, L)
        mod = capi.Modulator(M, K, L, taps, lib=lib)
        rx = capi.Demodulator(M, K, L, np.conj(taps), lib=lib)
        mod.set_stream(stream.cuda_stream)
        rx.set_stream(stream.cuda_stream)
        d_s = torch.from_numpy(crand(rng, frames, N)).cuda()
        d_x = torch.empty_like(d_s)
        d_y = torch.empty_like(d_s)
        ms = timed(lambda: mod.modulate_ptr(d_x.data_ptr(), d_s.data_ptr(), frames))
        report('modulator_kernel_cc::generic_work', 'K=%d M=%d L=%d' % (K, M, L), frames, ms, 16 * N, mod.last_kernel())
        ms = timed(lambda: rx.demodulate_ptr(d_y.data_ptr(), d_x.data_ptr(), 0, frames))
        report('receiver_kernel_cc::generic_work', 'K=%d M=%d L=%d' % (K, M, L), frames, ms, 16 * N, rx.last_kernel())
        del d_s, d_x, d_y


if __name__ == '__main__':
    what = sys.argv[1:] or ['tx']
    if 'shapes' in what:
        shape_rows()
    if 'tx' in what:
        tx_chain(9, 64, 52, 16, 8, 1 << 16)            # BASELINE configs[1]
        tx_chain(15, 1024, 832, 64, 32, 4096)           # headline shape with CP
        tx_chain(15, 256, 208, 32, 16, 1 << 14, (0, 16))
    if 'rx' in what:
        rx_rows(15, 256, 208, 1 << 14)                  # BASELINE configs[3]: estimator + advanced receiver, 4 SIC iterations
        rx_rows(15, 1024, 832, 4096)
        rx_eq_row(15, 2048, 2048)
    if 'rx2048' in what:
        rx_eq_row(15, 2048, 2048)
    if 'copy' in what:
        copy_rows(9, 64, 52, 16, 8, 1 << 16)
        copy_rows(15, 1024, 832, 64, 32, 4096)
    if 'next' in what:
        next_rows(15, 1024, 4096)
