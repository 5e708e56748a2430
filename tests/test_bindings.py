"""pybind11 module `gfdm_python` (gr-gfdm_b200/python/bindings.cc) and the C++ host layer
(include/gfdm_b200.hpp, include/gfdm/*.h).

CPU part: the module imports, carries the reference's class/method names
(python/bindings/*_python.cc), validates constructor arguments with the reference's
exception types (std::invalid_argument -> ValueError, std::runtime_error -> RuntimeError)
and never computes without a GPU.  GPU part (-m gpu): the checks of the reference's
python/qa_python_bindings.py:66-529 restated against the CPU oracle.
"""
import os
import subprocess

import numpy as np
import pytest

from conftest import PKG, ROOT, assert_complex_close
from gfdm_b200 import capi, design

REF_SURFACE = {  # class -> methods bound by the reference
    'Modulator': ['block_size', 'filter_taps', 'modulate'],
    'Demodulator': ['timeslots', 'subcarriers', 'overlap', 'block_size', 'filter_taps', 'demodulate',
                    'fft_filter_downsample', 'transform_subcarriers_to_td', 'demodulate_equalize',
                    'fft_equalize_filter_downsample', 'cancel_sc_interference'],
    'Cyclic_prefixer': ['block_size', 'frame_size', 'cyclic_shift', 'add_cyclic_prefix', 'remove_cyclic_prefix'],
    'Resource_mapper': ['block_size', 'frame_size', 'map_to_resources', 'demap_from_resources'],
    'Preamble_channel_estimator': ['timeslots', 'subcarriers', 'active_subcarriers', 'frame_len', 'is_dc_free',
                                   'estimate_frame', 'estimate_snr'],
}
NEW_SURFACE = {
    'Modulator': ['modulate_batch'], 'Demodulator': ['demodulate_batch', 'ic_filter_taps'],
    'Cyclic_prefixer': ['add_cyclic_prefix_batch', 'remove_cyclic_prefix_batch'],
    'Resource_mapper': ['map_to_resources_batch', 'demap_from_resources_batch'],
    'Preamble_channel_estimator': ['estimate_frame_batch', 'estimate_snr_cnrs', 'preamble_filter_taps'],
    'Advanced_receiver': ['block_size', 'set_ic', 'get_ic', 'set_phase_compensation', 'get_phase_compensation',
                          'demodulate', 'demodulate_equalize', 'demodulate_batch'],
    'Transmitter': ['input_vector_size', 'output_vector_size', 'cyclic_shifts', 'generic_work', 'generic_work_batch',
                    'generic_work_all_batch'],
}


@pytest.fixture(scope='module')
def gp():
    try:
        from gfdm_b200 import gfdm_python
    except ImportError:
        subprocess.run(['make', '-C', PKG, 'all'], check=True, stdout=subprocess.DEVNULL)
        from gfdm_b200 import gfdm_python
    return gfdm_python


def test_module_surface(gp):
    assert gp.backend() == 'cuda-sm_100a'
    for table in (REF_SURFACE, NEW_SURFACE):
        for cls, methods in table.items():
            for m in methods:
                assert hasattr(getattr(gp, cls), m), '%s.%s missing' % (cls, m)


def test_exception_types_match_the_reference(gp):
    taps = design.get_frequency_domain_filter('rrc', .5, 5, 16, 2)
    with pytest.raises(ValueError, match=r'number of frequency taps\(10\) MUST be equal to n_timeslots\(6\)'):
        gp.Modulator(6, 16, 2, taps)
    with pytest.raises(ValueError, match='overlap MUST be greater or equal 2'):
        gp.Demodulator(10, 16, 1, taps)
    with pytest.raises(ValueError, match='MUST be unique'):
        gp.Resource_mapper(5, 32, 4, [1, 2, 2, 3], True)
    with pytest.raises(ValueError, match='number of window taps'):
        gp.Cyclic_prefixer(80, 4, 2, 2, np.ones(5, np.complex64))
    if gp.device_count() == 0:  # no CPU fallback: a valid ctor fails loudly without a GPU
        with pytest.raises(RuntimeError, match='no usable CUDA device'):
            gp.Modulator(5, 16, 2, taps)


CPP_PROBE = r'''
#include <gfdm/add_cyclic_prefix_cc.h>
#include <gfdm/advanced_receiver_kernel_cc.h>
#include <gfdm/modulator_kernel_cc.h>
#include <gfdm/preamble_channel_estimator_cc.h>
#include <gfdm/receiver_kernel_cc.h>
#include <gfdm/resource_mapper_kernel_cc.h>
#include <gfdm/transmitter_kernel.h>
#include <cstdio>
#include <cstring>
using namespace gr::gfdm;
typedef std::complex<float> cf;
int main(int argc, char** argv)
{
    int caught = 0;
    try { modulator_kernel_cc m(6, 16, 2, std::vector<cf>(10, cf(1, 0))); } catch (const std::invalid_argument&) { ++caught; }
    try { receiver_kernel_cc r(5, 16, 1, std::vector<cf>(5, cf(1, 0))); } catch (const std::invalid_argument&) { ++caught; }
    try { resource_mapper_kernel_cc r(5, 32, 4, { 1, 2, 2, 3 }, true); } catch (const std::invalid_argument&) { ++caught; }
    try { remove_prefix r(100, 90, 11); } catch (const std::invalid_argument&) { ++caught; }
    try { extract_burst e(0, 0); } catch (const std::invalid_argument&) { ++caught; }
    try { symbol_mapper s(constellation{ std::vector<cf>(5, cf(1, 0)), GFDM_DECISION_QPSK_SIGN }); } catch (const std::invalid_argument&) { ++caught; }
    if (caught != 6) { printf("FAIL exceptions %d\n", caught); return 1; }
    if (argc < 4) { printf("OK exceptions\n"); return 0; }
    // round trip on the device: argv[1] taps (M*L cf), argv[2] symbols (N cf) -> argv[3] soft symbols
    const int M = 5, K = 16, L = 2, N = M * K;
    std::vector<cf> taps(M * L), rx(M * L), d(N), x(N), y(N);
    FILE* f = fopen(argv[1], "rb"); if (!f || fread(taps.data(), sizeof(cf), taps.size(), f) != taps.size()) return 2; fclose(f);
    f = fopen(argv[2], "rb"); if (!f || fread(d.data(), sizeof(cf), d.size(), f) != d.size()) return 2; fclose(f);
    for (size_t i = 0; i < taps.size(); ++i) rx[i] = std::conj(taps[i]);
    modulator_kernel_cc mod(M, K, L, taps);
    receiver_kernel_cc dem(M, K, L, rx);
    mod.generic_work(x.data(), d.data());
    dem.generic_work(y.data(), x.data());
    // legacy 2-D interface gives the same answer
    receiver_kernel_cc::matrix fd(K, std::vector<cf>(M)), td(K, std::vector<cf>(M));
    dem.filter_superposition(fd, x.data());
    dem.demodulate_subcarrier(td, fd);
    std::vector<cf> y2(N);
    dem.serialize_output(y2.data(), td);
    double err = 0, nrm = 0;   // the 2-D path runs the stages separately: same result up to fp32 rounding
    for (int i = 0; i < N; ++i) { err += std::norm(y[i] - y2[i]); nrm += std::norm(y[i]); }
    if (!(err <= 1e-10 * nrm)) { printf("FAIL legacy 2-D path differs (%g)\n", err / nrm); return 3; }
    // rows either side of the path: chunks -> modulate == lookup -> modulate; decisions == deciding the soft symbols;
    // remove_prefix / extract_burst copy what they should
    {
        symbol_mapper sm(constellation::qpsk());
        std::vector<unsigned char> ch(N), dec(N), dec2(N);
        for (int i = 0; i < N; ++i) ch[i] = (unsigned char)((i * 7 + 3) % 4);
        std::vector<cf> s1(N), x1(N), x2(N), soft(N);
        sm.map_chunks(s1.data(), ch.data(), N);
        mod.generic_work(x1.data(), s1.data());
        sm.modulate_chunks(mod, x2.data(), ch.data(), 1);
        if (memcmp(x1.data(), x2.data(), sizeof(cf) * N)) { printf("FAIL modulate_chunks\n"); return 4; }
        sm.demodulate_decide(dem, dec.data(), x2.data(), nullptr, 1);
        dem.generic_work(soft.data(), x2.data());
        sm.decide(dec2.data(), soft.data(), N);
        if (memcmp(dec.data(), dec2.data(), N)) { printf("FAIL demodulate_decide\n"); return 4; }
        remove_prefix rp(N + 12, N, 8);
        std::vector<cf> fr(2 * (N + 12)), blk(2 * N);
        for (size_t i = 0; i < fr.size(); ++i) fr[i] = cf((float)i, -(float)i);
        rp.work_batch(blk.data(), fr.data(), 2);
        if (blk[0] != fr[8] || blk[N] != fr[N + 12 + 8] || blk[2 * N - 1] != fr[N + 12 + 8 + N - 1]) { printf("FAIL remove_prefix\n"); return 4; }
        extract_burst eb(16, 2);
        std::vector<cf> bursts(2 * 16);
        extract_burst::result r = eb.work(bursts.data(), 2, fr.data(), (long long)fr.size(), { 1, 40 }, { 2.0f, 0.5f });
        if (r.n_produced != 2 || r.n_consumed != 56 || bursts[0] != cf(0, 0) || bursts[1] != fr[0] * 2.0f || bursts[16] != fr[38] * 0.5f) {
            printf("FAIL extract_burst %d %lld\n", r.n_produced, r.n_consumed); return 4; }
    }
    f = fopen(argv[3], "wb"); fwrite(y.data(), sizeof(cf), y.size(), f); fclose(f);
    printf("OK %d launches\n", (int)(mod.launch_count() + dem.launch_count()));
    return 0;
}
'''


@pytest.fixture(scope='module')
def cpp_probe(tmp_path_factory):
    d = tmp_path_factory.mktemp('cpp')
    src = d / 'probe.cc'
    src.write_text(CPP_PROBE)
    exe = d / 'probe'
    lib = os.path.join(PKG, 'lib')
    if not os.path.exists(os.path.join(lib, 'libgfdm_b200.so')):
        subprocess.run(['make', '-C', PKG, 'all'], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(['g++', '-std=c++17', '-Wall', '-I' + os.path.join(ROOT, 'include'), str(src), '-o', str(exe),
                    '-L' + lib, '-lgfdm_b200', '-Wl,-rpath,' + lib], check=True)
    return str(exe), d


def test_cpp_layer_compiles_with_reference_include_names(cpp_probe):
    exe, _ = cpp_probe
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and 'OK exceptions' in out.stdout, out.stdout + out.stderr


# ---------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_cpp_layer_round_trip(cpp_probe, port):
    exe, d = cpp_probe
    M, K, L = 5, 16, 2
    taps = design.get_frequency_domain_filter('rrc', .5, M, K, L).astype(np.complex64)
    sym = design.get_random_qpsk(M * K, rng=np.random.RandomState(5)).astype(np.complex64)
    taps.tofile(d / 'taps.bin')
    sym.tofile(d / 'sym.bin')
    out = subprocess.run([exe, str(d / 'taps.bin'), str(d / 'sym.bin'), str(d / 'y.bin')], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.startswith('OK'), out.stdout + out.stderr
    y = np.fromfile(d / 'y.bin', np.complex64)
    ref = capi.Demodulator(M, K, L, np.conj(taps), lib=port).demodulate(capi.Modulator(M, K, L, taps, lib=port).modulate(sym))
    assert_complex_close(y, ref, what='C++ modulator->receiver round trip')


@pytest.mark.gpu
@pytest.mark.parametrize('M,K,L', [(5, 16, 2), (16, 4, 2), (21, 128, 2), (9, 64, 2)])
def test_modulator_demodulator_like_qa_python_bindings(gp, port, M, K, L):
    """python/qa_python_bindings.py:66-240 -- modulate / demodulate / stages, complex128 input is cast."""
    taps = design.get_frequency_domain_filter('rrc', .5, M, K, L)
    d = design.get_random_qpsk(M * K, rng=np.random.RandomState(M + K))  # complex128 on purpose (forcecast)
    mod, dem = gp.Modulator(M, K, L, taps), gp.Demodulator(M, K, L, np.conj(taps))
    omod, odem = capi.Modulator(M, K, L, taps, lib=port), capi.Demodulator(M, K, L, np.conj(taps), lib=port)
    assert mod.block_size() == M * K and (dem.timeslots(), dem.subcarriers(), dem.overlap()) == (M, K, L)
    assert np.abs(np.array(mod.filter_taps()) - omod.filter_taps()).max() < 3e-7
    x = mod.modulate(d)
    assert x.dtype == np.complex64 and x.shape == (M * K,)
    assert_complex_close(x, omod.modulate(d), what='modulate')
    assert_complex_close(dem.demodulate(x), odem.demodulate(x), what='demodulate')
    fd = dem.fft_filter_downsample(x)
    assert_complex_close(fd, odem.fft_filter_downsample(x), what='fft_filter_downsample')
    assert_complex_close(dem.transform_subcarriers_to_td(fd), odem.transform_subcarriers_to_td(fd), what='to_td')
    eq = np.ones(M * K, np.complex64) * (0.8 - 0.3j)
    assert_complex_close(dem.demodulate_equalize(x, eq), odem.demodulate_equalize(x, eq), what='demodulate_equalize')
    assert_complex_close(dem.fft_equalize_filter_downsample(x, eq), odem.fft_equalize_filter_downsample(x, eq),
                         what='fft_equalize_filter_downsample')
    assert_complex_close(dem.cancel_sc_interference(d, fd), odem.cancel_sc_interference(d.astype(np.complex64), fd),
                         what='cancel_sc_interference')
    xb = mod.modulate_batch(np.stack([d, -d, 1j * d]))
    assert xb.shape == (3, M * K) and np.array_equal(xb[0], x)
    assert_complex_close(dem.demodulate_batch(xb), odem.demodulate_batch(xb), what='demodulate_batch')
    with pytest.raises(RuntimeError, match='Only ONE-dimensional vectors allowed!'):
        mod.modulate(np.zeros((2, M * K)))
    with pytest.raises(RuntimeError, match=r'MUST be equal to Modulator.block_size\(%d\)' % (M * K)):
        mod.modulate(np.zeros(M * K + 1))


@pytest.mark.gpu
def test_mapper_prefixer_estimator_transmitter_bindings(gp, port):
    """qa_python_bindings.py:242-529 (Cyclic_prefixer, Resource_mapper, Preamble_channel_estimator) + Transmitter."""
    cfg = design.get_gfdm_configuration()
    rng = np.random.RandomState(3)
    d = design.get_random_qpsk(cfg.timeslots * cfg.active_subcarriers, rng=rng)
    mp, omp = gp.Resource_mapper(cfg.timeslots, cfg.subcarriers, cfg.active_subcarriers, list(cfg.subcarrier_map), True), \
        capi.Resource_mapper(cfg.timeslots, cfg.subcarriers, cfg.active_subcarriers, cfg.subcarrier_map, True, lib=port)
    grid = mp.map_to_resources(d)
    assert np.array_equal(grid, omp.map_to_resources(d)), 'mapping is bit-exact'
    assert np.array_equal(mp.demap_from_resources(grid), d.astype(np.complex64))
    pf = gp.Cyclic_prefixer(cfg.block_len, cfg.cp_len, cfg.cs_len, cfg.ramp_len, cfg.window_taps)
    opf = capi.Cyclic_prefixer(cfg.block_len, cfg.cp_len, cfg.cs_len, cfg.ramp_len, cfg.window_taps, lib=port)
    x = design.get_random_qpsk(cfg.block_len, rng=rng)
    fr = pf.add_cyclic_prefix(x)
    assert fr.shape == (pf.frame_size(),) and np.array_equal(fr, opf.add_cyclic_prefix(x))
    assert np.array_equal(pf.remove_cyclic_prefix(fr), fr[cfg.cp_len:cfg.cp_len + cfg.block_len])
    est = gp.Preamble_channel_estimator(cfg.timeslots, cfg.subcarriers, cfg.active_subcarriers, True, 1, cfg.core_preamble)
    oest = capi.Preamble_channel_estimator(cfg.timeslots, cfg.subcarriers, cfg.active_subcarriers, True, 1,
                                           cfg.core_preamble, lib=port)
    h = est.estimate_frame(cfg.core_preamble)
    assert h.shape == (est.frame_len(),)
    assert np.abs(h - 1).max() < 1e-5, 'flat channel gives all ones (qa_channel_estimator_cc.py:63-86)'
    rxp = cfg.core_preamble * (0.5 + 0.2j) + 0.01 * design.get_random_qpsk(2 * cfg.subcarriers, rng=rng)
    assert_complex_close(est.estimate_frame(rxp), oest.estimate_frame(rxp), what='estimate_frame')
    assert abs(est.estimate_snr(rxp) - oest.estimate_snr(rxp)) <= 1e-3 * abs(oest.estimate_snr(rxp))
    tx = gp.Transmitter(cfg.timeslots, cfg.subcarriers, cfg.active_subcarriers, cfg.cp_len, cfg.cs_len, cfg.ramp_len,
                        list(cfg.subcarrier_map), True, cfg.overlap, cfg.tx_filter_taps, cfg.window_taps,
                        list(cfg.cyclic_shifts), [np.asarray(p) for p in cfg.full_preambles])
    otx = capi.Transmitter(cfg.timeslots, cfg.subcarriers, cfg.active_subcarriers, cfg.cp_len, cfg.cs_len, cfg.ramp_len,
                           cfg.subcarrier_map, True, cfg.overlap, cfg.tx_filter_taps, cfg.window_taps,
                           cfg.cyclic_shifts, cfg.full_preambles, lib=port)
    assert tx.input_vector_size() == d.size and tx.output_vector_size() == otx.output_vector_size()
    assert_complex_close(tx.generic_work(d), otx.generic_work(d), what='transmitter generic_work')
    db = np.stack([d, -d])
    assert_complex_close(tx.generic_work_batch(db), otx.work_batch(db), what='transmitter batch')


@pytest.mark.gpu
def test_advanced_receiver_binding(gp, port):
    M, K, L = 9, 64, 2
    taps = design.get_frequency_domain_filter('rrc', .5, M, K, L)
    smap = list(range(4, 52))
    rng = np.random.RandomState(11)
    sym = np.zeros((K, M), np.complex64)
    sym[smap] = design.get_random_qpsk(len(smap) * M, rng=rng).reshape(len(smap), M)
    x = capi.Modulator(M, K, L, taps, lib=port).modulate(sym.ravel())
    ar = gp.Advanced_receiver(M, K, L, np.conj(taps), smap, 3)
    oar = capi.Advanced_receiver(M, K, L, np.conj(taps), smap, 3, lib=port)
    assert ar.get_ic() == 3 and ar.block_size() == M * K
    assert_complex_close(ar.demodulate(x), oar.demodulate(x), what='advanced receiver')
    ar.set_ic(1)
    oar.set_ic(1)
    assert_complex_close(ar.demodulate_batch(np.stack([x, x])), oar.demodulate_batch(np.stack([x, x])), what='adv batch')


def test_pybind_new_rows_validate_before_device_use(gp):
    with pytest.raises(ValueError, match='MUST NOT exceed frame_len'):
        gp.Remove_prefix(100, 90, 11)
    with pytest.raises(ValueError, match='burst_len MUST be positive'):
        gp.Extract_burst(0, 0)
    with pytest.raises(ValueError, match='QPSK sign rule needs exactly 4'):
        gp.Symbol_mapper([1, -1], 1)


@pytest.mark.gpu
def test_pybind_rows_either_side_of_the_path(gp, port):
    """Symbol_mapper / Remove_prefix / Extract_burst of the pybind11 module against the oracle (ctypes)."""
    rng = np.random.default_rng(21)
    pts = design.qam16_points().astype(np.complex64)
    sm, so = gp.Symbol_mapper(pts, 0), capi.Symbol_mapper((pts, 0), lib=port)
    assert sm.n_points() == 16 and sm.bits_per_symbol() == 4
    chunks = rng.integers(0, 16, 999).astype(np.uint8)
    sym = sm.map_chunks(chunks)
    assert np.array_equal(sym, so.map_chunks(chunks))
    noisy = sym + 0.2 * (rng.standard_normal(999) + 1j * rng.standard_normal(999)).astype(np.complex64)
    assert np.array_equal(sm.decide(noisy), so.decide(noisy))
    bits = rng.integers(0, 2, 4 * 64).astype(np.uint8)
    assert np.array_equal(sm.bits2symbols(bits), so.bits2symbols(bits))
    assert np.array_equal(sm.symbols2bits(noisy[:64]), so.symbols2bits(noisy[:64]))
    M, K, L = 9, 64, 2
    taps = design.get_frequency_domain_filter('rrc', .5, M, K, L)
    mod, rx = gp.Modulator(M, K, L, taps), gp.Demodulator(M, K, L, np.conj(taps))
    ch = rng.integers(0, 16, (5, M * K)).astype(np.uint8)
    x = sm.modulate_chunks_batch(mod, ch)
    assert np.array_equal(x, mod.modulate_batch(sm.map_chunks(ch.ravel()).reshape(5, -1)))
    dec = sm.demodulate_decide_batch(rx, x)
    assert np.array_equal(dec, sm.decide(rx.demodulate_batch(x).ravel()).reshape(5, -1))
    frames = (rng.standard_normal((4, 100)) + 1j * rng.standard_normal((4, 100))).astype(np.complex64)
    assert np.array_equal(gp.Remove_prefix(100, 80, 12).work_batch(frames), frames[:, 12:92])
    stream = frames.ravel()
    bursts, produced, consumed = gp.Extract_burst(32, 4).work(stream, [2, 100, 390], [2.0, 1.0, 0.5])
    ob, oc = capi.Extract_burst(32, 4, lib=port).work(stream, [2, 100, 390], [2.0, 1.0, 0.5])
    assert produced == ob.shape[0] == 2 and consumed == oc and np.array_equal(bursts[:produced], ob)
