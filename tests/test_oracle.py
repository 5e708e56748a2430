"""Pin the oracles: (1) both CPU oracles against the golden vectors produced by the
reference's NumPy spec (pygfdm), (2) the plain-C restatement (oracle/gfdm_oracle.c)
against the unmodified reference C++ sources (oracle/_ref).  Tolerances are those
of the reference's own tests (python/qa_python_bindings.py) or tighter."""
import numpy as np
import pytest

from conftest import rel_l2
from gfdm_b200 import capi, design

MODDEMOD_KEYS = ['M16_K4_L2', 'M21_K128_L2', 'M5_K16_L2', 'M9_K64_L2', 'M15_K64_L2', 'M127_K16_L4',
                 'M5_K32_L2', 'M8_K16_L2', 'M3_K32_L2', 'M7_K8_L2']


def _mkl(key):
    return [int(s[1:]) for s in key.split('_')]


@pytest.fixture(params=['port', 'ref'])
def lib(request):
    return request.getfixturevalue(request.param)


@pytest.mark.parametrize('key', MODDEMOD_KEYS)
def test_modulator_vs_pygfdm(lib, golden, key):
    g = golden.moddemod
    M, K, L = _mkl(key)
    mod = capi.Modulator(M, K, L, g[key + '_taps'], lib=lib)
    assert mod.block_size() == M * K
    x = mod.modulate(g[key + '_data'])
    assert rel_l2(x, g[key + '_tx']) < 2e-6
    # places=5 of qa_python_bindings.py:254-294
    assert np.abs(x - g[key + '_tx']).max() < 0.5e-5


@pytest.mark.parametrize('key', [k for k in MODDEMOD_KEYS if k.endswith('L2')])
def test_demodulator_vs_pygfdm(lib, golden, key):
    g = golden.moddemod
    M, K, L = _mkl(key)
    dem = capi.Demodulator(M, K, L, g[key + '_taps'], lib=lib)
    assert (dem.timeslots(), dem.subcarriers(), dem.overlap(), dem.block_size()) == (M, K, L, M * K)
    assert np.abs(dem.filter_taps() - g[key + '_taps']).max() < 1e-6  # qa_python_bindings.py:304-319
    y = dem.demodulate(g[key + '_tx'])
    assert rel_l2(y, g[key + '_rx']) < 2e-6
    assert np.abs(y - g[key + '_rx']).max() < 0.5e-5
    y2 = dem.demodulate(g[key + '_rnd'])
    assert rel_l2(y2, g[key + '_rnd_rx']) < 2e-6
    # equalize with a constant e^{j} channel (qa_python_bindings.py:365-386)
    eq = np.ones(M * K) * np.exp(1j)
    y3 = dem.demodulate_equalize(g[key + '_tx'] * np.exp(1j), eq)
    assert rel_l2(y3, g[key + '_rx']) < 3e-6
    # stage-wise == fused (qa_python_bindings.py:388-440)
    fd = dem.fft_filter_downsample(g[key + '_tx'])
    assert rel_l2(dem.transform_subcarriers_to_td(fd), g[key + '_rx']) < 2e-6
    fd2 = dem.fft_equalize_filter_downsample(g[key + '_tx'] * np.exp(1j), eq)
    assert rel_l2(fd2, fd) < 3e-6


def test_sic_with_true_symbols_recovers_data(lib, golden):
    """qa_python_bindings.py:388-415 (test_005_steps)."""
    g = golden.moddemod
    key = 'M5_K32_L2'
    M, K, L = _mkl(key)
    dem = capi.Demodulator(M, K, L, g[key + '_taps'], lib=lib)
    fd = dem.fft_filter_downsample(g[key + '_tx'])
    data = g[key + '_data']
    for _ in range(2):
        res = dem.transform_subcarriers_to_td(dem.cancel_sc_interference(data, fd))
    assert np.abs(res - data).max() < 0.05  # places=1


@pytest.mark.parametrize('name', ['t001', 't002', 't003', 't004', 't005'])
def test_mapper_vs_pygfdm_bit_exact(lib, golden, name):
    g = golden.mapper
    M, K, A, per_ts = [int(v) for v in g[name + '_cfg']]
    mp = capi.Resource_mapper(M, K, A, g[name + '_map'], bool(per_ts), lib=lib)
    assert (mp.block_size(), mp.frame_size()) == (M * A, M * K)
    f = mp.map_to_resources(g[name + '_in'])
    assert np.array_equal(f, g[name + '_out'].astype(np.complex64))
    d = mp.demap_from_resources(g[name + '_out'])
    assert np.array_equal(d, g[name + '_in'].astype(np.complex64))


@pytest.mark.parametrize('name', ['p001', 'p002', 'p003', 'p004'])
def test_cyclic_prefix_vs_pygfdm(lib, golden, name):
    g = golden.cyclic_prefix
    N, cp, cs, ramp, shift = [int(v) for v in g[name + '_cfg']]
    pf = capi.Cyclic_prefixer(N, cp, cs, ramp, g[name + '_window'], shift, lib=lib)
    assert (pf.block_size(), pf.frame_size(), pf.cyclic_shift()) == (N, N + cp + cs, shift)
    res = pf.add_cyclic_prefix(g[name + '_in'])
    ref = g[name + '_out'].astype(np.complex64)
    assert np.abs(res - ref).max() < 1e-6
    # un-windowed samples are pure copies -> bit exact
    assert np.array_equal(res[ramp:-ramp], ref[ramp:-ramp])
    # the short (2*ramp_len) window form is equivalent (add_cyclic_prefix_cc.cc:42-56)
    w = g[name + '_window']
    pf2 = capi.Cyclic_prefixer(N, cp, cs, ramp, np.concatenate((w[:ramp], w[-ramp:])), shift, lib=lib)
    assert np.array_equal(pf2.add_cyclic_prefix(g[name + '_in']), res)
    frame = np.arange(N + cp + cs, dtype=np.complex64)
    assert np.array_equal(pf.remove_cyclic_prefix(frame), frame[cp:cp + N])


@pytest.mark.parametrize('name', ['e001', 'e002', 'e003'])
def test_estimator_vs_numpy_estimator(lib, golden, name):
    """C++ estimator vs. the independent NumPy one (pygfdm/validation_utils.py:33-78)."""
    g = golden.estimator
    M, K, A = [int(v) for v in g[name + '_cfg']]
    est = capi.Preamble_channel_estimator(M, K, A, True, 1, g[name + '_core'], lib=lib)
    assert (est.timeslots(), est.subcarriers(), est.active_subcarriers(), est.frame_len(),
            est.is_dc_free()) == (M, K, A, M * K, True)
    res = est.estimate_frame(g[name + '_rx'])
    ref = g[name + '_est']
    assert rel_l2(res, ref) < 5e-6
    # vs. the true channel on active bins (qa_python_bindings.py:452-490, places=1)
    fh = np.fft.fft(g[name + '_h'], M * K)
    n = M * A // 2
    assert np.abs(res[:n] - fh[:n]).max() < 0.05 and np.abs(res[-n:] - fh[-n:]).max() < 0.05
    taps = est.preamble_filter_taps()
    assert abs(taps.sum() - 1) < 1e-6 and taps.argmax() == 4


def test_estimator_flat_channel_is_one(lib, golden):
    """qa_channel_estimator_cc.py:63-86: undistorted preamble -> 1+0j everywhere (places=6)."""
    g = golden.preamble
    core = g['qpsk_32_24_core']
    est = capi.Preamble_channel_estimator(3, 32, 24, True, 1, core, lib=lib)
    res = est.estimate_frame(core)
    assert np.abs(res - 1.0).max() < 0.5e-6 * 4


def test_estimator_snr(lib, golden):
    """qa_python_bindings.py:492-529 (within 1 dB at 4 dB)."""
    core = golden.preamble['qpsk_1024_936_core']
    K, A = 1024, 936
    rng = np.random.default_rng(5)
    snr_lin = 10. ** (4.0 / 10.)
    energy = np.sum(np.abs(core) ** 2)
    nscale = 1. / np.sqrt(snr_lin) * np.sqrt((K / A) * 2. * energy / core.size)
    noise = rng.standard_normal(core.size) + 1j * rng.standard_normal(core.size)
    noise = noise / np.abs(noise) * nscale
    est = capi.Preamble_channel_estimator(5, K, A, True, 1, core, lib=lib)
    res = est.estimate_snr(core + noise)
    assert abs(10. * np.log10(res) - 4.0) < 1.0
    snr2, cnrs = est.estimate_snr_cnrs(core + noise)
    assert snr2 == res and cnrs.shape == (A,) and np.all(cnrs >= 0)


def test_transmitter_vs_pygfdm(lib, golden):
    """qa_transmitter_cc.py:80-183 (places=5), all four cyclic shifts."""
    g = golden.transmitter
    M, K, A, L, cp, cs, ramp = [int(v) for v in g['cfg']]
    shifts = [int(s) for s in g['shifts']]
    tx = capi.Transmitter(M, K, A, cp, cs, ramp, g['map'], True, L, g['taps'], g['window'], shifts,
                          list(g['preambles']), lib=lib)
    assert tx.input_vector_size() == M * A
    assert tx.output_vector_size() == g['preambles'].shape[1] + cp + M * K + cs
    assert tx.cyclic_shifts() == shifts
    data, frames = g['data'], g['frames']
    out = tx.work_all_batch(data)
    assert out.shape == frames.shape
    assert np.abs(out - frames).max() < 0.5e-5
    assert np.abs(tx.work_batch(data) - frames[0]).max() < 0.5e-5
    assert np.abs(tx.generic_work(data[1]) - frames[0][1]).max() < 0.5e-5
    blk = tx.modulate(data[0])
    assert np.abs(tx.add_frame(blk, 7) - frames[2][0]).max() < 0.5e-5


# ---- constructor validation: same conditions and messages as the reference ------------
def test_ctor_errors(lib):
    taps = design.get_frequency_domain_filter('rrc', .5, 5, 16, 2)
    with pytest.raises(ValueError, match=r'number of frequency taps\(10\) MUST be equal to n_timeslots\(6\) \* overlap\(2\) = 12!'):
        capi.Modulator(6, 16, 2, taps, lib=lib)
    with pytest.raises(ValueError, match=r'number of frequency taps\(10\) MUST be equal'):
        capi.Demodulator(6, 16, 2, taps, lib=lib)
    with pytest.raises(ValueError, match='overlap MUST be greater or equal 2'):
        capi.Demodulator(10, 16, 1, taps, lib=lib)
    with pytest.raises(ValueError, match=r'active_subcarriers\(40\) MUST be smaller or equal to subcarriers\(32\)!'):
        capi.Resource_mapper(5, 32, 40, np.arange(40), lib=lib)
    with pytest.raises(ValueError, match=r'number of subcarrier_map entries\(3\) MUST be equal to active_subcarriers\(4\)!'):
        capi.Resource_mapper(5, 32, 4, [1, 2, 3], lib=lib)
    with pytest.raises(ValueError, match='MUST be unique'):
        capi.Resource_mapper(5, 32, 4, [1, 2, 2, 3], lib=lib)
    with pytest.raises(ValueError, match='greater or equal to ZERO'):
        capi.Resource_mapper(5, 32, 4, [-1, 2, 4, 3], lib=lib)
    with pytest.raises(ValueError, match='smaller or equal to subcarriers'):
        capi.Resource_mapper(5, 32, 4, [1, 2, 4, 33], lib=lib)
    with pytest.raises(ValueError, match=r'number of window taps\(7\) MUST be equal to 2\*ramp_len\(8\) OR block_len\+cp_len \(120\)!'):
        capi.Cyclic_prefixer(96, 16, 8, 4, np.ones(7), lib=lib)
    cfg = design.get_gfdm_configuration()
    args = (cfg.timeslots, cfg.subcarriers, cfg.active_subcarriers, cfg.cp_len, cfg.cs_len, cfg.ramp_len,
            cfg.subcarrier_map, True, cfg.overlap, cfg.tx_filter_taps, cfg.window_taps)
    with pytest.raises(ValueError, match='Number of cyclic shifts and number of preambles do not match!'):
        capi.Transmitter(*args, [0, 1], cfg.full_preambles, lib=lib)
    with pytest.raises(ValueError, match='All preambles must have equal size!'):
        capi.Transmitter(*args, [0, 1], [cfg.full_preambles[0], cfg.full_preambles[0][:-1]], lib=lib)


def test_mapper_size_errors_and_zero_padding(lib):
    mp = capi.Resource_mapper(3, 8, 4, [1, 2, 6, 7], True, lib=lib)
    with pytest.raises(ValueError, match=r'input vector size\(13\) MUST not exceed active_subcarriers \* timeslots\(12\)!'):
        mp.map_to_resources_n(np.ones(13), 13)
    with pytest.raises(ValueError, match=r'output vector size\(13\) MUST not exceed'):
        mp.demap_from_resources_n(np.ones(24), 13)
    d = np.arange(12, dtype=np.complex64) + 1
    short = mp.map_to_resources_n(d, 7)   # inputs beyond ninput_size are replaced by 0
    full = mp.map_to_resources(np.concatenate((d[:7], np.zeros(5))))
    assert np.array_equal(short, full)
    assert mp.map_to_resources_n(d, 0).any() == False  # empty input -> all-zero grid


def test_demap_per_subcarrier_quirk(lib):
    """lib/resource_mapper_kernel_cc.cc:155-159 writes element [noutput_size] too."""
    mp = capi.Resource_mapper(3, 8, 4, [1, 2, 6, 7], False, lib=lib)
    grid = np.arange(24, dtype=np.complex64) + 1
    out = mp.demap_from_resources_n(grid, 5, pad=1)
    assert np.array_equal(out[:5], grid[[3, 4, 5, 6, 7]])
    assert out[5] == grid[8]
    # the batch entry keeps that write inside the frame
    got = mp.demap_from_resources_batch(np.stack([grid, grid + 100]), 5)
    assert np.array_equal(got[0], grid[[3, 4, 5, 6, 7]]) and np.array_equal(got[1], grid[[3, 4, 5, 6, 7]] + 100)


# ---- restatement vs. the unmodified reference sources ----------------------------------
@pytest.mark.parametrize('M,K,L', [(5, 16, 2), (9, 64, 2), (15, 256, 2), (15, 1024, 2), (127, 16, 4), (6, 12, 3), (4, 10, 2)])
def test_port_vs_ref_mod_demod(port, ref, M, K, L):
    rng = np.random.default_rng(M * 1000 + K)
    taps = design.get_frequency_domain_filter('rrc', .5, M, K, L) if (M * K) % 2 == 0 or True else None
    taps = taps * (1.3 - 0.2j)  # non-normalised, complex: exercises the ctor renormalisation
    d = design.get_random_qam16(M * K, rng)
    mp, mr = capi.Modulator(M, K, L, taps, lib=port), capi.Modulator(M, K, L, taps, lib=ref)
    assert np.abs(mp.filter_taps() - mr.filter_taps()).max() < 2e-7
    xp, xr = mp.modulate(d), mr.modulate(d)
    assert rel_l2(xp, xr) < 1e-6
    rx = np.conj(taps)
    dp, dr = capi.Demodulator(M, K, L, rx, lib=port), capi.Demodulator(M, K, L, rx, lib=ref)
    assert np.abs(dp.ic_filter_taps() - dr.ic_filter_taps()).max() < 4e-7
    assert rel_l2(dp.demodulate(xr), dr.demodulate(xr)) < 1e-6
    eq = (rng.standard_normal(M * K) + 1j * rng.standard_normal(M * K)) * .3 + 1
    assert rel_l2(dp.demodulate_equalize(xr, eq), dr.demodulate_equalize(xr, eq)) < 1e-6
    fd = dr.fft_filter_downsample(xr)
    assert rel_l2(dp.cancel_sc_interference(d, fd), dr.cancel_sc_interference(d, fd)) < 1e-6


@pytest.mark.parametrize('rule,phase_comp,ic', [(capi.DECISION_QPSK_SIGN, 0, 4), (capi.DECISION_QPSK_SIGN, 1, 2),
                                                  (capi.DECISION_NEAREST, 0, 3), (capi.DECISION_NEAREST, 1, 1),
                                                  (capi.DECISION_QPSK_SIGN, 0, 0)])
def test_port_vs_ref_advanced_receiver(port, ref, rule, phase_comp, ic):
    M, K, L, A = 9, 32, 2, 20
    rng = np.random.default_rng(77 + ic)
    taps = design.get_frequency_domain_filter('rrc', .5, M, K, L)
    smap = design.get_subcarrier_map(K, A)
    if rule == capi.DECISION_QPSK_SIGN:
        pts, _ = capi.qpsk_constellation()
    else:
        pts = design.qam16_points().astype(np.complex64)
    syms = pts[rng.integers(0, pts.size, M * A)]
    grid = capi.Resource_mapper(M, K, A, smap, True, lib=ref).map_to_resources(syms)
    x = capi.Modulator(M, K, L, taps, lib=ref).modulate(grid)
    x = x * np.exp(0.05j) + 0.002 * (rng.standard_normal(M * K) + 1j * rng.standard_normal(M * K))
    eq = np.full(M * K, 1.02 * np.exp(0.01j))
    outs = []
    for lib in (port, ref):
        rx = capi.Advanced_receiver(M, K, L, np.conj(taps), smap, ic, (pts, rule), phase_comp, lib=lib)
        assert rx.get_ic() == ic and rx.get_phase_compensation() == phase_comp and rx.block_size() == M * K
        outs.append((rx.demodulate(x), rx.demodulate_equalize(x, eq)))
    assert rel_l2(outs[0][0], outs[1][0]) < 2e-6
    assert rel_l2(outs[0][1], outs[1][1]) < 2e-6
    if ic == 0:  # qa_advanced_receiver_sb_cc.py:45-80: no iterations == simple receiver
        simple = capi.Demodulator(M, K, L, np.conj(taps), lib=ref).demodulate(x)
        assert np.array_equal(outs[1][0], simple)


def test_advanced_receiver_recovers_qpsk(lib):
    """qa_advanced_receiver_sb_cc.py:82-119: alpha=1.0 RRC, 64 SIC iterations recover the symbols (places=2)."""
    M, K, L = 9, 64, 2
    taps = design.get_frequency_domain_filter('rrc', 1.0, M, K, L)
    pts, rule = capi.qpsk_constellation()
    rng = np.random.default_rng(9)
    d = pts[rng.integers(0, 4, M * K)]
    x = capi.Modulator(M, K, L, taps, lib=lib).modulate(d)
    rx = capi.Advanced_receiver(M, K, L, taps, np.arange(K), 64, (pts, rule), 0, lib=lib)
    assert np.abs(rx.demodulate(x) - d).max() < 0.5e-2
    rx.set_ic(2)
    assert rx.get_ic() == 2


@pytest.mark.parametrize('dc_free', [True, False])
def test_port_vs_ref_estimator_steps(port, ref, golden, dc_free):
    g = golden.estimator
    M, K, A = [int(v) for v in g['e003_cfg']]
    core, rx = g['e003_core'], g['e003_rx']
    res = []
    rng = np.random.default_rng(11)
    noisy = (rx + 0.01 * (rng.standard_normal(rx.size) + 1j * rng.standard_normal(rx.size))).astype(np.complex64)
    for lib in (port, ref):
        est = capi.Preamble_channel_estimator(M, K, A, dc_free, 1, core, lib=lib)
        h = est.estimate_preamble_channel(rx)
        f = est.filter_preamble_estimate(h)
        fr = est.interpolate_frame(f, fill=7.0)   # 7.0 marks bins the reference never writes
        res.append((h, f, fr, est.estimate_frame(rx, fill=7.0), est.prepare_for_zf(fr), est.estimate_snr_cnrs(noisy)))
    # unused preamble bins hold 0.5 / ~0 garbage in the reference: compare active bins only
    act = design.get_subcarrier_map(K, A, dc_free=True)
    assert rel_l2(res[0][0][act], res[1][0][act]) < 2e-6
    for a, b in zip(res[0][1:5], res[1][1:5]):
        assert rel_l2(a, b) < 2e-6
    assert np.array_equal(res[0][2] == 7.0, res[1][2] == 7.0)
    if not dc_free:  # M bins stay unwritten
        assert int((res[1][2] == 7.0).sum()) == M
    assert abs(res[0][5][0] - res[1][5][0]) <= 1e-4 * abs(res[1][5][0])
    assert rel_l2(res[0][5][1], res[1][5][1]) < 1e-4


def test_port_vs_ref_batches(port, ref):
    """frame f uses in + f*in_size, out + f*out_size; eq advances per frame."""
    M, K, L = 5, 16, 2
    rng = np.random.default_rng(3)
    taps = design.get_frequency_domain_filter('rrc', .5, M, K, L)
    d = (rng.standard_normal((6, M * K)) + 1j * rng.standard_normal((6, M * K))).astype(np.complex64)
    eq = (1 + .2 * rng.standard_normal((6, M * K)) + .2j * rng.standard_normal((6, M * K))).astype(np.complex64)
    for lib in (port, ref):
        mod, dem = capi.Modulator(M, K, L, taps, lib=lib), capi.Demodulator(M, K, L, taps, lib=lib)
        xb = mod.modulate_batch(d)
        assert np.array_equal(xb, np.stack([mod.modulate(f) for f in d]))
        assert np.array_equal(dem.demodulate_batch(xb), np.stack([dem.demodulate(f) for f in xb]))
        assert np.array_equal(dem.demodulate_batch(xb, eq),
                              np.stack([dem.demodulate_equalize(f, e) for f, e in zip(xb, eq)]))
        with pytest.raises(RuntimeError, match='GFDM_MEM_DEVICE'):
            mod.modulate_ptr(0, 0, 1)


def test_signal_energy_and_fft(port, ref):
    rng = np.random.default_rng(1)
    for n in [1, 2, 5, 15, 16, 80, 127, 576, 3840]:
        x = (rng.standard_normal((3, n)) + 1j * rng.standard_normal((3, n))).astype(np.complex64)
        for lib in (port, ref):
            assert rel_l2(capi.FFT(n, True, lib=lib).execute(x), np.fft.fft(x.astype(np.complex128), axis=1)) < 1e-6
            assert rel_l2(capi.FFT(n, False, lib=lib).execute(x), np.fft.ifft(x.astype(np.complex128), axis=1) * n) < 1e-6
            assert abs(lib.calculate_signal_energy(x[0]) - np.sum(np.abs(x[0]) ** 2)) < 1e-3 * n


def test_estimator_rejects_odd_active_subcarriers(lib):
    """ADVICE r1 / deliberate deviation: for odd active_subcarriers the reference's interpolate_frame writes up to
    index N + M/2 - 1 of its N-long frame (dc free), filter_preamble_estimate correlates against an unfilled entry and
    estimate_snr leaves cnrs[A-1] unset.  Every backend of the ABI refuses the configuration at create."""
    core = np.ones(64, np.complex64)
    with pytest.raises(ValueError, match='active_subcarriers MUST be even'):
        capi.Preamble_channel_estimator(5, 32, 21, True, 1, core, lib=lib)


def test_advanced_receiver_rejects_inconsistent_constellations(lib):
    """the sign rule indexes points[0..3]; an unknown rule value is refused as well."""
    taps = design.get_frequency_domain_filter('rrc', .5, 5, 16, 2)
    pts = np.array([1, -1], np.complex64)
    with pytest.raises(ValueError, match='exactly 4 constellation points'):
        capi.Advanced_receiver(5, 16, 2, taps, np.arange(16), 1, (pts, capi.DECISION_QPSK_SIGN), 0, lib=lib)
    with pytest.raises(ValueError, match='unknown constellation decision rule'):
        capi.Advanced_receiver(5, 16, 2, taps, np.arange(16), 1, (pts, 7), 0, lib=lib)
