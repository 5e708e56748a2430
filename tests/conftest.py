"""Shared fixtures.  `-m "not gpu"` tests run on CPU only (oracles, host logic, ABI
surface); `-m gpu` tests are the parity tests proper and call the CUDA library
through the C ABI."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'gr-gfdm_b200')
if PKG not in sys.path:
    sys.path.insert(0, PKG)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault('HOME', '/tmp')  # the reference reads $HOME unchecked (gfdm_kernel_utils.cc:37)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')
ORACLE_DIR = os.path.join(ROOT, 'oracle')
PORT_SO = os.path.join(ORACLE_DIR, '_ref', 'libgfdm_port.so')
REF_SO = os.path.join(ORACLE_DIR, '_ref', 'libgfdm_ref.so')
PRODUCT_SO = os.path.join(PKG, 'lib', 'libgfdm_b200.so')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def _build_oracles():
    if not (os.path.exists(PORT_SO) and (os.path.exists(REF_SO) or not os.path.isdir('/root/reference/lib'))):
        subprocess.run(['make', '-C', ORACLE_DIR, 'all'], check=True, stdout=subprocess.DEVNULL)


@pytest.fixture(scope='session')
def port():
    """oracle/gfdm_oracle.c -- the plain-C restatement (always available)."""
    from gfdm_b200 import capi
    _build_oracles()
    return capi.load(PORT_SO)


@pytest.fixture(scope='session')
def ref():
    """oracle/_ref/libgfdm_ref.so -- the unmodified reference sources + shims."""
    from gfdm_b200 import capi
    _build_oracles()
    if not os.path.exists(REF_SO):
        pytest.skip('reference oracle not built (no /root/reference and no prebuilt .so)')
    return capi.load(REF_SO)


@pytest.fixture(scope='session', params=['port', 'ref'])
def oracle(request):
    """Both checkers in turn: the plain-C restatement and the unmodified reference sources (+ shims).  The GPU parity
    tests take this fixture, so every CUDA result is compared with the REFERENCE's own code as well."""
    return request.getfixturevalue(request.param)


@pytest.fixture(scope='session')
def cuda():
    """The product library; gpu tests fail (not skip) when it cannot be loaded."""
    from gfdm_b200 import capi
    lib = capi.load()
    assert lib.backend().startswith('cuda'), lib.backend()
    assert lib.device_count() > 0, 'no CUDA device visible'
    return lib


@pytest.fixture(scope='session')
def golden():
    class G(object):
        def __getattr__(self, name):
            return np.load(os.path.join(GOLDEN, name + '.npz'))
    return G()


def rel_l2(a, b):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def max_abs_over_rms(a, b):
    a = np.asarray(a).ravel()
    b = np.asarray(b).ravel()
    rms = np.sqrt(np.mean(np.abs(b) ** 2))
    return float(np.max(np.abs(a - b)) / max(rms, 1e-30))


def assert_complex_close(res, ref, rel=1e-5, mx=1e-4, what=''):
    """north_star tolerance: rel-L2 <= 1e-5 and max-abs <= 1e-4 of signal RMS."""
    r, m = rel_l2(res, ref), max_abs_over_rms(res, ref)
    assert r <= rel and m <= mx, '%s rel-l2 %.3e (<= %.1e), max-abs/rms %.3e (<= %.1e)' % (what, r, rel, m, mx)


# (M, K) with a single-kernel (shared-memory resident) modulator / receiver: csrc/fused_shapes_*.cu, parameters chosen by
# tools/shape_chooser.py -- every power-of-two K in 16..1024 for M in {3, 5, 7, 9, 15, 21} (M = 21 does not fit at K = 1024)
# plus the reference's other test shapes; K = 2048 / M = 15 runs the two-pass kernels.
FUSED_TABLE = set((M, K) for K in (16, 32, 64, 128, 256, 512, 1024) for M in (3, 5, 7, 9, 15, 21)) - {(21, 1024)} | \
    {(8, 16), (16, 4), (7, 8), (19, 32), (16, 64)}


def fused_expected(M, K):
    return (M, K) in FUSED_TABLE or (M, K) == (15, 2048)
