"""CPU-side checks of the product library: it loads, exports every symbol that
include/gfdm_b200.h declares, validates constructor arguments exactly like the
reference *before* touching the GPU, and fails loudly (no CPU fallback) when no
CUDA device is present."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import PRODUCT_SO, ROOT
from gfdm_b200 import capi, design


@pytest.fixture(scope='module')
def product():
    if not os.path.exists(PRODUCT_SO):
        import __graft_entry__ as g
        g.build()
    return capi.load(PRODUCT_SO)


def _declared_symbols():
    hdr = open(os.path.join(ROOT, 'include', 'gfdm_b200.h')).read()
    return sorted(set(re.findall(r'GFDM_B200_API[^;(]*?\b(gfdm_[a-z0-9_]+)\s*\(', hdr)))


def test_header_and_binding_agree():
    assert _declared_symbols() == sorted(capi.Library._PROTOS)


def test_library_exports_every_declared_symbol(product):
    out = subprocess.run(['nm', '-D', '--defined-only', PRODUCT_SO], check=True, capture_output=True, text=True).stdout
    exported = set(l.split()[-1] for l in out.splitlines() if ' T ' in l)
    missing = [s for s in _declared_symbols() if s not in exported]
    assert not missing, missing
    assert product.backend() == 'cuda-sm_100a'


def test_oracles_export_the_same_abi(port, ref):
    for lib in (port, ref):
        assert lib.exported_symbols == sorted(capi.Library._PROTOS)
    assert port.backend() == 'oracle-port' and ref.backend() == 'oracle-ref'


def test_sass_is_sm100a(product):
    out = subprocess.run(['cuobjdump', '-lelf', PRODUCT_SO], capture_output=True, text=True).stdout
    assert 'sm_100a' in out, out[:500]


def test_ctor_validation_precedes_device_use(product):
    """Same conditions/messages as the reference; raised even without a GPU."""
    taps = design.get_frequency_domain_filter('rrc', .5, 5, 16, 2)
    with pytest.raises(ValueError, match=r'number of frequency taps\(10\) MUST be equal to n_timeslots\(6\) \* overlap\(2\) = 12!'):
        capi.Modulator(6, 16, 2, taps, lib=product)
    with pytest.raises(ValueError, match='overlap MUST be greater or equal 2'):
        capi.Demodulator(10, 16, 1, taps, lib=product)
    with pytest.raises(ValueError, match='MUST be unique'):
        capi.Resource_mapper(5, 32, 4, [1, 2, 2, 3], lib=product)
    with pytest.raises(ValueError, match=r'number of window taps\(7\)'):
        capi.Cyclic_prefixer(96, 16, 8, 4, np.ones(7), lib=product)
    cfg = design.get_gfdm_configuration()
    args = (cfg.timeslots, cfg.subcarriers, cfg.active_subcarriers, cfg.cp_len, cfg.cs_len, cfg.ramp_len,
            cfg.subcarrier_map, True, cfg.overlap, cfg.tx_filter_taps, cfg.window_taps)
    with pytest.raises(ValueError, match='Number of cyclic shifts and number of preambles do not match!'):
        capi.Transmitter(*args, [0, 1], cfg.full_preambles, lib=product)


def test_no_cpu_fallback(product):
    if product.device_count() > 0:
        pytest.skip('a GPU is present')
    taps = design.get_frequency_domain_filter('rrc', .5, 5, 16, 2)
    with pytest.raises(capi.GfdmCudaError, match='no usable CUDA device'):
        capi.Modulator(5, 16, 2, taps, lib=product)
    with pytest.raises(capi.GfdmCudaError, match='no CPU fallback'):
        capi.Resource_mapper(5, 32, 4, [1, 2, 3, 4], lib=product)


def test_product_never_references_the_oracle():
    pkg = os.path.join(ROOT, 'gr-gfdm_b200')
    for dirpath, _, files in os.walk(pkg):
        if 'build' in dirpath:
            continue
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h', '.cc', '.cpp')):
                txt = open(os.path.join(dirpath, f), errors='replace').read()
                assert 'libgfdm_port' not in txt and 'libgfdm_ref' not in txt and 'gfdm_oracle' not in txt, f


def test_fused_shape_table_matches_the_sources():
    """The set of single-kernel shapes the GPU tests expect (conftest.FUSED_TABLE) is exactly what the translation units
    csrc/fused_shapes_*.cu instantiate (GFDM_SHAPE(M, R1, R2, ...): K = R1 * R2), and what tools/shape_chooser.py accepts."""
    import glob
    import importlib.util
    from conftest import FUSED_TABLE
    found = set()
    for path in glob.glob(os.path.join(ROOT, 'gr-gfdm_b200', 'csrc', 'fused_shapes_*.cu')):
        for m in re.finditer(r'^\s*GFDM_SHAPE\((\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+)', open(path).read(), re.M):
            M, R1, R2, T, IPT, MINB = (int(x) for x in m.groups())
            assert (M, R1 * R2) not in found, 'shape instantiated twice: M=%d K=%d' % (M, R1 * R2)
            found.add((M, R1 * R2))
    assert found == set(FUSED_TABLE), (sorted(found - set(FUSED_TABLE)), sorted(set(FUSED_TABLE) - found))
    spec = importlib.util.spec_from_file_location('shape_chooser', os.path.join(ROOT, 'tools', 'shape_chooser.py'))
    chooser = importlib.util.module_from_spec(spec)
    src = open(os.path.join(ROOT, 'tools', 'shape_chooser.py')).read()
    assert 'def shape(' in src
    # the chooser's budget arithmetic accepts every instantiated parameter set (it mirrors Shape's static_asserts)
    exec(compile(src.split("if __name__")[0], 'shape_chooser', 'exec'), chooser.__dict__)
    for path in glob.glob(os.path.join(ROOT, 'gr-gfdm_b200', 'csrc', 'fused_shapes_k*.cu')):   # (the baseline unit is hand-tuned)
        for m in re.finditer(r'^\s*GFDM_SHAPE\((\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+),\s*(\d+)', open(path).read(), re.M):
            M, R1, R2, T, IPT, MINB = (int(x) for x in m.groups())
            assert chooser.shape(M, R1, R2, T, IPT, MINB) is not None, (M, R1, R2, T, IPT, MINB)
