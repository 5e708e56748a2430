"""GNU Radio block shims and the legacy 2-D receiver interface.

include/gfdm_b200_blocks.hpp is compiled against tests/stub_gnuradio/ (a stand-in for the GNU Radio headers the blocks
use: GNU Radio cannot be installed in this image) and tests/cpp/block_shims_probe.cc plays the scheduler: every block's
work() over several frames must equal the kernel class called frame by frame, which is the loop the reference blocks run
(lib/simple_modulator_cc_impl.cc:72-76, advanced_receiver_sb_cc_impl.cc:98-113, transmitter_cc_impl.cc:165-177, ...).

The 2-D interface is compared with the REFERENCE's own vector<vector<>> methods (oracle/ref_capi.cc exports a test-only
entry that runs them): on the CPU the plain-C oracle behind the same C++ adapters, on the GPU the product."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import PKG, PORT_SO, ROOT, assert_complex_close
from gfdm_b200 import capi, design

LIBDIR = os.path.join(PKG, 'lib')


def _compile(tmp, src, exe, extra_inc=(), lib='gfdm_b200', libdir=LIBDIR):
    if not os.path.exists(os.path.join(LIBDIR, 'libgfdm_b200.so')):
        subprocess.run(['make', '-C', PKG, 'all'], check=True, stdout=subprocess.DEVNULL)
    cmd = ['g++', '-std=c++17', '-O1', '-Wall', '-I' + os.path.join(ROOT, 'include')] + ['-I' + i for i in extra_inc] + \
          [os.path.join(ROOT, 'tests', 'cpp', src), '-o', str(tmp / exe), '-L' + libdir, '-l' + lib, '-Wl,-rpath,' + libdir, '-lpthread']
    subprocess.run(cmd, check=True)
    return str(tmp / exe)


@pytest.fixture(scope='module')
def blocks_probe(tmp_path_factory):
    d = tmp_path_factory.mktemp('blocks')
    return _compile(d, 'block_shims_probe.cc', 'blocks', extra_inc=[os.path.join(ROOT, 'tests', 'stub_gnuradio')])


def test_block_shims_compile_and_validate_without_a_device(blocks_probe):
    out = subprocess.run([blocks_probe, '--dry'], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and 'OK dry' in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_block_shims_work_equals_per_frame_kernel_calls(blocks_probe):
    out = subprocess.run([blocks_probe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and 'OK blocks' in out.stdout, out.stdout + out.stderr


def _legacy_inputs(tmp, M, K, L):
    rng = np.random.default_rng(M + K)
    taps = np.conj(design.get_frequency_domain_filter('rrc', .5, M, K, L)).astype(np.complex64)
    x = (rng.standard_normal(M * K) + 1j * rng.standard_normal(M * K)).astype(np.complex64)
    taps.tofile(str(tmp / 'taps.bin'))
    x.tofile(str(tmp / 'x.bin'))
    return taps, x


def _reference_2d(ref, M, K, L, taps, x):
    """filter_superposition / demodulate_subcarrier / remove_sc_interference run by the reference's own 2-D methods."""
    dem = capi.Demodulator(M, K, L, taps, lib=ref)
    fn = ref.dll.gfdm_ref_receiver_legacy_2d
    fn.restype = ctypes.c_int
    fn.argtypes = [ctypes.c_void_p] * 5
    out = [np.empty(M * K, np.complex64) for _ in range(3)]
    assert fn(dem._h, *[o.ctypes.data_as(ctypes.c_void_p) for o in out], x.ctypes.data_as(ctypes.c_void_p)) == 0
    return out


def _probe_2d(exe, tmp, M, K, L):
    out = subprocess.run([exe, str(M), str(K), str(L), str(tmp / 'taps.bin'), str(tmp / 'x.bin'), str(tmp / 'out.bin')],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and 'OK' in out.stdout, out.stdout + out.stderr
    return np.fromfile(str(tmp / 'out.bin'), np.complex64).reshape(3, M * K)


@pytest.mark.parametrize('M,K,L', [(5, 16, 2), (9, 64, 2)])
def test_legacy_2d_adapters_on_the_port_oracle_vs_reference_2d(tmp_path, ref, M, K, L):
    """The C++ adapters themselves (include/gfdm_b200.hpp) on a CPU library exporting the ABI."""
    exe = _compile(tmp_path, 'legacy2d_probe.cc', 'legacy2d_port', lib='gfdm_port', libdir=os.path.dirname(PORT_SO))
    taps, x = _legacy_inputs(tmp_path, M, K, L)
    got = _probe_2d(exe, tmp_path, M, K, L)
    for g, w, what in zip(got, _reference_2d(ref, M, K, L, taps, x), ('filter_superposition', 'demodulate_subcarrier', 'remove_sc_interference')):
        assert_complex_close(g, w, what=what)


@pytest.mark.gpu
@pytest.mark.parametrize('M,K,L', [(5, 16, 2), (9, 64, 2), (15, 256, 2), (21, 128, 2)])
def test_legacy_2d_interface_on_the_gpu_vs_reference_2d(tmp_path, ref, M, K, L):
    exe = _compile(tmp_path, 'legacy2d_probe.cc', 'legacy2d_gpu')
    taps, x = _legacy_inputs(tmp_path, M, K, L)
    got = _probe_2d(exe, tmp_path, M, K, L)
    for g, w, what in zip(got, _reference_2d(ref, M, K, L, taps, x), ('filter_superposition', 'demodulate_subcarrier', 'remove_sc_interference')):
        assert_complex_close(g, w, what=what)


@pytest.mark.gpu
def test_single_process_multi_gpu_driver_is_bit_identical(tmp_path):
    """gr::gfdm::multi_gpu: one host thread per device (per-thread device selection), contiguous shards, no exchange;
    output == the single-handle output bit for bit.  With one GPU visible the workers share it (still separate handles,
    streams and threads); `gpurun --gpus N` exercises N devices with the same binary."""
    exe = _compile(tmp_path, 'multi_gpu_probe.cc', 'multi_gpu')
    for workers, frames in ((2, 257), (3, 40), (8, 1000)):
        out = subprocess.run([exe, str(workers), str(frames)], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and out.stdout.startswith('OK'), out.stdout + out.stderr


def test_multi_gpu_driver_host_logic_without_a_device(tmp_path):
    """The worker-pool logic of gr::gfdm::multi_gpu with a stand-in kernel type: a factory that throws fails the constructor
    (workers joined, nothing left waiting -- a hang here cost a GPU session), job exceptions reach the caller once, shards
    partition the batch."""
    src = tmp_path / 'mg_logic.cc'
    src.write_text(r'''
#include <gfdm_b200.hpp>
#include <atomic>
#include <cstdio>
using namespace gr::gfdm;
struct fake { int id; };
int main()
{
    bool threw = false;
    try {
        multi_gpu<fake> bad({ 0, 1 }, [&]() -> std::unique_ptr<fake> { throw std::runtime_error("no device"); });
    } catch (const std::exception&) { threw = true; }
    if (!threw) { printf("FAIL ctor\n"); return 1; }
    for (size_t n : { (size_t)0, (size_t)1, (size_t)7, (size_t)4096 })
        for (size_t w : { (size_t)1, (size_t)3, (size_t)8 }) {
            size_t next = 0;
            for (size_t r = 0; r < w; ++r) {
                const auto b = multi_gpu<fake>::shard_bounds(n, w, r);
                if (b.first != next || b.second < b.first) { printf("FAIL bounds\n"); return 1; }
                next = b.second;
            }
            if (next != n) { printf("FAIL bounds cover\n"); return 1; }
        }
    printf("OK\n");
    return 0;
}
''')
    exe = tmp_path / 'mg_logic'
    subprocess.run(['g++', '-std=c++17', '-I' + os.path.join(ROOT, 'include'), str(src), '-o', str(exe), '-L' + LIBDIR, '-lgfdm_b200',
                    '-Wl,-rpath,' + LIBDIR, '-lpthread'], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and 'OK' in out.stdout, out.stdout + out.stderr
