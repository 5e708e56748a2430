#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ from the reference's own
NumPy specification (``pygfdm``, /root/reference/python/pygfdm).

Runs ONLY in the build container (the GPU box has no /root/reference); the
resulting ``*.npz`` files are committed next to this script.  Every expected
value is produced by unmodified pygfdm code; the only additions are
 * aliases for NumPy/SciPy names pygfdm still uses (np.complex, np.float,
   np.int, scipy.signal.gaussian), and
 * a stub ``commpy`` module (scikit-commpy is not installed, no network) that
   provides the published closed-form ``rrcosfilter`` / ``rcosfilter``.
The shapes follow the reference's known-answer tests
(python/qa_python_bindings.py:66-529, qa_transmitter_cc.py:80-183,
qa_channel_estimator_cc.py:63-86, qa_simple_modulator_cc.py:69-97).

usage:  python tests/golden/make_golden.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/python'


def install_shims():
    import scipy.signal
    import scipy.signal.windows
    np.complex = complex
    np.float = float
    np.int = int
    scipy.signal.gaussian = scipy.signal.windows.gaussian

    commpy = types.ModuleType('commpy')

    def rrcosfilter(N, alpha, Ts, Fs):
        # published closed form (scikit-commpy filters.rrcosfilter)
        N = int(N)
        T_delta = 1.0 / float(Fs)
        time_idx = (np.arange(N) - N / 2) * T_delta
        h = np.zeros(N, dtype=float)
        for x in range(N):
            t = (x - N / 2) * T_delta
            if t == 0.0:
                h[x] = 1.0 - alpha + (4 * alpha / np.pi)
            elif alpha != 0 and (t == Ts / (4 * alpha) or t == -Ts / (4 * alpha)):
                h[x] = (alpha / np.sqrt(2)) * (((1 + 2 / np.pi) * (np.sin(np.pi / (4 * alpha)))) +
                                               ((1 - 2 / np.pi) * (np.cos(np.pi / (4 * alpha)))))
            else:
                h[x] = (np.sin(np.pi * t * (1 - alpha) / Ts) +
                        4 * alpha * (t / Ts) * np.cos(np.pi * t * (1 + alpha) / Ts)) / \
                       (np.pi * t * (1 - (4 * alpha * t / Ts) * (4 * alpha * t / Ts)) / Ts)
        return time_idx, h

    def rcosfilter(N, alpha, Ts, Fs):
        N = int(N)
        T_delta = 1.0 / float(Fs)
        time_idx = (np.arange(N) - N / 2) * T_delta
        h = np.zeros(N, dtype=float)
        for x in range(N):
            t = (x - N / 2) * T_delta
            if t == 0.0:
                h[x] = 1.0
            elif alpha != 0 and (t == Ts / (2 * alpha) or t == -Ts / (2 * alpha)):
                h[x] = (np.pi / 4) * (np.sin(np.pi * t / Ts) / (np.pi * t / Ts))
            else:
                h[x] = (np.sin(np.pi * t / Ts) / (np.pi * t / Ts)) * \
                       (np.cos(np.pi * alpha * t / Ts) / (1 - (((2 * alpha * t) / Ts) * ((2 * alpha * t) / Ts))))
        return time_idx, h

    commpy.rrcosfilter = rrcosfilter
    commpy.rcosfilter = rcosfilter
    sys.modules['commpy'] = commpy
    sys.dont_write_bytecode = True
    sys.path.insert(0, REF)


def c128(x):
    return np.asarray(x, dtype=np.complex128)


def main():
    install_shims()
    from pygfdm.filters import get_frequency_domain_filter
    from pygfdm.gfdm_modulation import gfdm_modulate_block
    from pygfdm.gfdm_receiver import gfdm_demodulate_block
    from pygfdm.mapping import (get_data_matrix, get_subcarrier_map,
                                map_to_waveform_resources)
    from pygfdm.cyclic_prefix import (add_cyclic_starfix, get_raised_cosine_ramp,
                                      get_root_raised_cosine_ramp, get_window_len,
                                      pinch_block)
    from pygfdm.preamble import mapped_preamble
    from pygfdm.utils import get_random_qpsk, get_random_samples
    from pygfdm.validation_utils import frame_estimator
    from pygfdm.zadoff_chu import generate_zadoff_chu_sequence

    out = {}

    # ---- filter taps (python/pygfdm/filters.py:27-54) -----------------------
    taps = {}
    for ft, alpha, M, K, L in [('rrc', .35, 16, 4, 2), ('rrc', .35, 21, 128, 2),
                               ('rrc', .5, 5, 16, 2), ('rrc', .2, 9, 64, 2),
                               ('rrc', .5, 9, 64, 2), ('rrc', .5, 15, 64, 2),
                               ('rrc', .5, 127, 16, 4), ('rrc', .5, 127, 16, 2),
                               ('rrc', 1.0, 9, 64, 2), ('rrc', .5, 9, 32, 2),
                               ('rrc', .35, 5, 32, 2), ('rrc', .35, 25, 96, 2),
                               ('rc', .5, 8, 16, 2), ('rrc', .5, 15, 256, 2),
                               ('rrc', .5, 3, 32, 2), ('rrc', .5, 7, 8, 2)]:
        taps['%s_%g_%d_%d_%d' % (ft, alpha, M, K, L)] = c128(
            get_frequency_domain_filter(ft, alpha, M, K, L))
    np.savez_compressed(os.path.join(HERE, 'taps.npz'), **taps)

    # ---- modulator / demodulator (qa_python_bindings.py:254-440) ------------
    moddemod = {}
    cases = [(16, 4, 2, .35, 101), (21, 128, 2, .35, 102), (5, 16, 2, .5, 103),
             (9, 64, 2, .2, 104), (15, 64, 2, .5, 105), (127, 16, 4, .5, 106),
             (5, 32, 2, .35, 107), (8, 16, 2, .5, 108), (3, 32, 2, .5, 109),
             (7, 8, 2, .5, 110)]
    for M, K, L, alpha, seed in cases:
        key = 'M%d_K%d_L%d' % (M, K, L)
        t = get_frequency_domain_filter('rrc', alpha, M, K, L)
        d = get_random_qpsk(M * K, seed)
        D = get_data_matrix(d, K, group_by_subcarrier=False)
        x = gfdm_modulate_block(D, t, M, K, L, False)
        moddemod[key + '_taps'] = c128(t)
        moddemod[key + '_data'] = c128(d)
        moddemod[key + '_tx'] = c128(x)
        if L == 2:  # gfdm_extract_subcarriers hard-codes L=2 (gfdm_receiver.py:54)
            y = gfdm_demodulate_block(x, t, K, M, L)
            moddemod[key + '_rx'] = c128(y)
            # demodulating an arbitrary (non-GFDM) input, qa_advanced_receiver_sb_cc.py:45-80
            r = get_random_samples(M * K, seed + 1000)
            moddemod[key + '_rnd'] = c128(r)
            moddemod[key + '_rnd_rx'] = c128(gfdm_demodulate_block(r, t, K, M, L))
    np.savez_compressed(os.path.join(HERE, 'moddemod.npz'), **moddemod)

    # ---- resource mapper (qa_python_bindings.py:66-165) ---------------------
    mp = {}
    for name, M, K, A, smap, per_ts in [
            ('t001', 15, 32, 24, np.arange(4, 28), True),
            ('t002', 15, 32, 24, np.arange(4, 28), False),
            ('t003', 15, 32, 24, get_subcarrier_map(32, 24, True), True),
            ('t004', 9, 64, 52, get_subcarrier_map(64, 52, True), False),
            ('t005', 9, 32, 20, get_subcarrier_map(32, 20, False), True)]:
        d = np.arange(M * A, dtype=np.complex64) + 1
        f = map_to_waveform_resources(d, A, K, smap, per_ts)
        mp[name + '_cfg'] = np.array([M, K, A, int(per_ts)])
        mp[name + '_map'] = np.asarray(smap, dtype=np.int64)
        mp[name + '_in'] = c128(d)
        mp[name + '_out'] = c128(f)
    np.savez_compressed(os.path.join(HERE, 'mapper.npz'), **mp)

    # ---- cyclic prefix (qa_python_bindings.py:175-244) ----------------------
    cp = {}
    for name, M, K, cpl, csl, rl, shift, root in [('p001', 19, 32, 16, 8, 4, 0, False),
                                                   ('p002', 3, 32, 16, 8, 4, 4, False),
                                                   ('p003', 9, 64, 16, 8, 8, 0, False),
                                                   ('p004', 8, 16, 8, 4, 4, 3, True)]:
        N = M * K
        wl = get_window_len(cpl, M, K, csl)
        w = (get_root_raised_cosine_ramp if root else get_raised_cosine_ramp)(rl, wl)
        data = get_random_samples(N, 200 + M)
        ref = np.concatenate((data[-(cpl + shift):], data, data[0:csl - shift]))
        ref2 = add_cyclic_starfix(np.roll(data, shift), cpl, csl)  # qa_transmitter_cc.py:49-53
        assert np.all(ref == ref2)
        ref = pinch_block(ref, w)
        cp[name + '_cfg'] = np.array([N, cpl, csl, rl, shift])
        cp[name + '_window'] = c128(w)
        cp[name + '_in'] = c128(data)
        cp[name + '_out'] = c128(ref)
    np.savez_compressed(os.path.join(HERE, 'cyclic_prefix.npz'), **cp)

    # ---- preambles (python/pygfdm/preamble.py:91-132) -----------------------
    pr = {}
    for name, seed, alpha, A, K, cpl, rl, zc, shift in [
            ('zc_64_52', 4711, .5, 52, 64, 16, 8, True, 0),
            ('zc_64_52_s3', 4711, .5, 52, 64, 16, 8, True, 3),
            ('qpsk_64_52', 3660365253, .5, 52, 64, 32, 16, False, 0),
            ('qpsk_32_24', 3660365253, .5, 24, 32, 16, 8, False, 0),
            ('zc_256_208', 3660365253, .5, 208, 256, 16, 8, True, 0),
            ('qpsk_1024_936', 3660365253, .5, 936, 1024, 512, 256, False, 0)]:
        smap = get_subcarrier_map(K, A, dc_free=True)
        full, core = mapped_preamble(seed, 'rrc', alpha, A, K, smap, 2, cpl, rl,
                                     use_zadoff_chu=zc, cyclic_shift=shift)
        pr[name + '_cfg'] = np.array([seed, A, K, cpl, rl, int(zc), shift], dtype=np.int64)
        pr[name + '_alpha'] = np.array([alpha])
        pr[name + '_full'] = c128(full)
        pr[name + '_core'] = c128(core)
    pr['zc_52_19'] = c128(generate_zadoff_chu_sequence(52, 19))
    np.savez_compressed(os.path.join(HERE, 'preamble.npz'), **pr)

    # ---- channel estimator vs. the independent NumPy estimator --------------
    # (python/pygfdm/validation_utils.py:33-78; qa_python_bindings.py:452-490)
    es = {}
    for name, M, K, A, cpl, rl, h in [
            ('e001', 5, 64, 52, 32, 16, np.array([1., .5, .1j, .1 + .05j])),
            ('e002', 3, 32, 24, 16, 8, np.array([1.])),
            ('e003', 15, 256, 208, 16, 8, np.array([.8, .3 - .2j, 0, .1j, 0, 0, .05]))]:
        smap = get_subcarrier_map(K, A, dc_free=True)
        full, core = mapped_preamble(3660365253, 'rrc', .5, A, K, smap, 2, cpl, rl)
        rx = np.convolve(full, h, 'full')[0:full.size][cpl:-rl]
        fe = frame_estimator(core, K, M, A)
        es[name + '_cfg'] = np.array([M, K, A])
        es[name + '_core'] = c128(core)
        es[name + '_rx'] = c128(rx)
        es[name + '_h'] = c128(h)
        es[name + '_est'] = c128(fe.estimate_frame(rx))
    np.savez_compressed(os.path.join(HERE, 'estimator.npz'), **es)

    # ---- full TX chain (qa_transmitter_cc.py:41-183) -------------------------
    tx = {}
    M, K, A, L = 9, 64, 52, 2
    cpl, csl = 16, 8
    shifts = [0, 3, 7, 8]
    t = get_frequency_domain_filter('rrc', .5, M, K, L)
    w = get_raised_cosine_ramp(csl, get_window_len(cpl, M, K, csl))
    smap = get_subcarrier_map(K, A, True)
    pre = [mapped_preamble(4711, 'rrc', .5, A, K, smap, L, cpl, csl,
                           use_zadoff_chu=True, cyclic_shift=s)[0] for s in shifts]
    n_frames = 3
    data = []
    ref = [[] for _ in shifts]
    for i in range(n_frames):
        d = get_random_qpsk(A * M, 300 + i)
        dd = map_to_waveform_resources(d, A, K, smap)
        D = get_data_matrix(dd, K, group_by_subcarrier=False)
        b = gfdm_modulate_block(D, t, M, K, L, False)
        for j, (s, p) in enumerate(zip(shifts, pre)):
            f = pinch_block(add_cyclic_starfix(np.roll(b, s), cpl, csl), w)
            ref[j].append(np.concatenate((p, f)))
        data.append(d)
    tx['cfg'] = np.array([M, K, A, L, cpl, csl, csl])
    tx['shifts'] = np.array(shifts)
    tx['taps'] = c128(t)
    tx['window'] = c128(w)
    tx['map'] = np.asarray(smap, dtype=np.int64)
    tx['preambles'] = c128(np.array(pre))
    tx['data'] = c128(np.array(data))
    tx['frames'] = c128(np.array(ref))  # [shift][frame][sample]
    np.savez_compressed(os.path.join(HERE, 'transmitter.npz'), **tx)

    for f in sorted(os.listdir(HERE)):
        if f.endswith('.npz'):
            print('%-22s %8d B' % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == '__main__':
    main()
