"""ctypes binding of the C ABI declared in include/gfdm_b200.h.

The classes mirror the reference's pybind11 module ``gfdm_python``
(python/bindings/*_python.cc: Modulator, Demodulator, Cyclic_prefixer,
Resource_mapper, Preamble_channel_estimator) -- same constructor arguments, same
method names, same size checks and messages -- and add ``Advanced_receiver`` and
``Transmitter`` plus ``*_batch`` methods (arrays shaped [n_frames, size]) and
``*_ptr`` methods that take raw device pointers (zero-copy, asynchronous).

``load(path)`` binds any shared library exporting the ABI.  The default is the
product library (hand-written sm_100a CUDA).  Tests load the CPU oracles through
the same code by passing their path explicitly; nothing in this package refers
to oracle/.
"""
import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_float, c_int, c_longlong, c_size_t, c_void_p

import numpy as np

GFDM_OK, GFDM_ERR_INVALID_ARGUMENT, GFDM_ERR_RUNTIME, GFDM_ERR_CUDA, GFDM_ERR_UNSUPPORTED = range(5)
MEM_HOST, MEM_DEVICE, MEM_HOST_ASYNC = 0, 1, 2
DECISION_NEAREST, DECISION_QPSK_SIGN = 0, 1

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIBRARY = os.path.join(os.path.dirname(_PKG_DIR), 'lib', 'libgfdm_b200.so')


class GfdmCudaError(RuntimeError):
    pass


class _Constellation(ctypes.Structure):
    _fields_ = [('points', c_void_p), ('n_points', c_int), ('decision_rule', c_int)]


def _c64(a):
    """py::array_t<complex<float>, c_style | forcecast> semantics."""
    return np.ascontiguousarray(a, dtype=np.complex64)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _ptr(a):
    return a.ctypes.data_as(c_void_p)


class Library(object):
    """One loaded shared library exporting include/gfdm_b200.h."""

    _PROTOS = {
        # name: (restype, [argtypes])
        'gfdm_last_error': (c_char_p, []),
        'gfdm_backend': (c_char_p, []),
        'gfdm_device_count': (c_int, []),
        'gfdm_set_device': (c_int, [c_int]),
        'gfdm_set_stream': (c_int, [c_void_p, c_void_p]),
        'gfdm_sync': (c_int, [c_void_p]),
        'gfdm_launch_count': (c_longlong, [c_void_p]),
        'gfdm_last_kernel': (c_char_p, [c_void_p]),
        'gfdm_calculate_signal_energy': (c_int, [POINTER(c_float), c_void_p, c_int]),
        'gfdm_fft_create': (c_int, [POINTER(c_void_p), c_int, c_int]),
        'gfdm_fft_destroy': (None, [c_void_p]),
        'gfdm_fft_execute_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int]),
        'gfdm_modulator_create': (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_void_p, c_int]),
        'gfdm_modulator_destroy': (None, [c_void_p]),
        'gfdm_modulator_block_size': (c_int, [c_void_p]),
        'gfdm_modulator_filter_taps': (c_int, [c_void_p, c_void_p]),
        'gfdm_modulator_work': (c_int, [c_void_p, c_void_p, c_void_p]),
        'gfdm_modulator_work_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int]),
        'gfdm_receiver_create': (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_void_p, c_int]),
        'gfdm_receiver_destroy': (None, [c_void_p]),
        'gfdm_receiver_block_size': (c_int, [c_void_p]),
        'gfdm_receiver_timeslots': (c_int, [c_void_p]),
        'gfdm_receiver_subcarriers': (c_int, [c_void_p]),
        'gfdm_receiver_overlap': (c_int, [c_void_p]),
        'gfdm_receiver_filter_taps': (c_int, [c_void_p, c_void_p]),
        'gfdm_receiver_ic_filter_taps': (c_int, [c_void_p, c_void_p]),
        'gfdm_receiver_work': (c_int, [c_void_p, c_void_p, c_void_p]),
        'gfdm_receiver_work_equalize': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
        'gfdm_receiver_work_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int]),
        'gfdm_receiver_fft_filter_downsample': (c_int, [c_void_p, c_void_p, c_void_p]),
        'gfdm_receiver_fft_equalize_filter_downsample': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
        'gfdm_receiver_fft_filter_downsample_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int]),
        'gfdm_receiver_transform_subcarriers_to_td': (c_int, [c_void_p, c_void_p, c_void_p]),
        'gfdm_receiver_transform_subcarriers_to_td_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int]),
        'gfdm_receiver_cancel_sc_interference': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
        'gfdm_receiver_cancel_sc_interference_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int]),
        'gfdm_advanced_receiver_create': (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_void_p, c_int,
                                                  c_void_p, c_int, c_int, POINTER(_Constellation), c_int]),
        'gfdm_advanced_receiver_destroy': (None, [c_void_p]),
        'gfdm_advanced_receiver_block_size': (c_int, [c_void_p]),
        'gfdm_advanced_receiver_set_ic': (c_int, [c_void_p, c_int]),
        'gfdm_advanced_receiver_get_ic': (c_int, [c_void_p]),
        'gfdm_advanced_receiver_set_phase_compensation': (c_int, [c_void_p, c_int]),
        'gfdm_advanced_receiver_get_phase_compensation': (c_int, [c_void_p]),
        'gfdm_advanced_receiver_work': (c_int, [c_void_p, c_void_p, c_void_p]),
        'gfdm_advanced_receiver_work_equalize': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
        'gfdm_advanced_receiver_work_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int]),
        'gfdm_resource_mapper_create': (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_void_p, c_int, c_int, c_int]),
        'gfdm_resource_mapper_destroy': (None, [c_void_p]),
        'gfdm_resource_mapper_frame_size': (c_size_t, [c_void_p]),
        'gfdm_resource_mapper_block_size': (c_size_t, [c_void_p]),
        'gfdm_resource_mapper_input_vector_size': (c_size_t, [c_void_p]),
        'gfdm_resource_mapper_output_vector_size': (c_size_t, [c_void_p]),
        'gfdm_resource_mapper_map_to_resources': (c_int, [c_void_p, c_void_p, c_void_p, c_size_t]),
        'gfdm_resource_mapper_demap_from_resources': (c_int, [c_void_p, c_void_p, c_void_p, c_size_t]),
        'gfdm_resource_mapper_map_to_resources_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int]),
        'gfdm_resource_mapper_demap_from_resources_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int]),
        'gfdm_cyclic_prefixer_create': (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_void_p, c_int, c_int]),
        'gfdm_cyclic_prefixer_destroy': (None, [c_void_p]),
        'gfdm_cyclic_prefixer_block_size': (c_int, [c_void_p]),
        'gfdm_cyclic_prefixer_frame_size': (c_int, [c_void_p]),
        'gfdm_cyclic_prefixer_cyclic_shift': (c_int, [c_void_p]),
        'gfdm_cyclic_prefixer_work': (c_int, [c_void_p, c_void_p, c_void_p]),
        'gfdm_cyclic_prefixer_add_cyclic_prefix': (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
        'gfdm_cyclic_prefixer_remove_cyclic_prefix': (c_int, [c_void_p, c_void_p, c_void_p]),
        'gfdm_cyclic_prefixer_add_cyclic_prefix_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int]),
        'gfdm_cyclic_prefixer_remove_cyclic_prefix_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int]),
        'gfdm_channel_estimator_create': (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_void_p, c_int]),
        'gfdm_channel_estimator_destroy': (None, [c_void_p]),
        'gfdm_channel_estimator_fft_len': (c_int, [c_void_p]),
        'gfdm_channel_estimator_timeslots': (c_int, [c_void_p]),
        'gfdm_channel_estimator_frame_len': (c_int, [c_void_p]),
        'gfdm_channel_estimator_active_subcarriers': (c_int, [c_void_p]),
        'gfdm_channel_estimator_is_dc_free': (c_int, [c_void_p]),
        'gfdm_channel_estimator_preamble_filter_taps': (c_int, [c_void_p, c_void_p]),
        'gfdm_channel_estimator_estimate_preamble_channel': (c_int, [c_void_p, c_void_p, c_void_p]),
        'gfdm_channel_estimator_filter_preamble_estimate': (c_int, [c_void_p, c_void_p, c_void_p]),
        'gfdm_channel_estimator_interpolate_frame': (c_int, [c_void_p, c_void_p, c_void_p]),
        'gfdm_channel_estimator_prepare_for_zf': (c_int, [c_void_p, c_void_p, c_void_p]),
        'gfdm_channel_estimator_estimate_frame': (c_int, [c_void_p, c_void_p, c_void_p]),
        'gfdm_channel_estimator_estimate_frame_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int]),
        'gfdm_channel_estimator_estimate_snr': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
        'gfdm_channel_estimator_estimate_snr_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int]),
        'gfdm_transmitter_create': (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int,
                                            c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_int,
                                            c_void_p, c_int, POINTER(c_void_p), c_void_p, c_int]),
        'gfdm_transmitter_destroy': (None, [c_void_p]),
        'gfdm_transmitter_input_vector_size': (c_int, [c_void_p]),
        'gfdm_transmitter_output_vector_size': (c_int, [c_void_p]),
        'gfdm_transmitter_n_cyclic_shifts': (c_int, [c_void_p]),
        'gfdm_transmitter_cyclic_shifts': (c_int, [c_void_p, c_void_p]),
        'gfdm_transmitter_work': (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
        'gfdm_transmitter_modulate': (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
        'gfdm_transmitter_add_frame': (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
        'gfdm_transmitter_work_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int]),
        'gfdm_transmitter_work_all_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int]),
        'gfdm_transmitter_set_chain_fusion': (c_int, [c_void_p, c_int]),
        # rows either side of the path
        'gfdm_remove_prefix_create': (c_int, [POINTER(c_void_p), c_int, c_int, c_int]),
        'gfdm_remove_prefix_destroy': (None, [c_void_p]),
        'gfdm_remove_prefix_work_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int]),
        'gfdm_extract_burst_create': (c_int, [POINTER(c_void_p), c_int, c_int, c_int]),
        'gfdm_extract_burst_destroy': (None, [c_void_p]),
        'gfdm_extract_burst_activate_cfo_compensation': (c_int, [c_void_p, c_int]),
        'gfdm_extract_burst_work': (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_longlong, c_void_p, c_void_p,
                                            c_void_p, c_int, POINTER(c_int), POINTER(c_longlong), c_int]),
        'gfdm_symbol_mapper_create': (c_int, [POINTER(c_void_p), POINTER(_Constellation)]),
        'gfdm_symbol_mapper_destroy': (None, [c_void_p]),
        'gfdm_symbol_mapper_n_points': (c_int, [c_void_p]),
        'gfdm_symbol_mapper_bits_per_symbol': (c_int, [c_void_p]),
        'gfdm_symbol_mapper_points': (c_int, [c_void_p, c_void_p]),
        'gfdm_symbol_mapper_decision_rule': (c_int, [c_void_p]),
        'gfdm_symbol_mapper_map_chunks_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int]),
        'gfdm_symbol_mapper_decide_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int]),
        'gfdm_symbol_mapper_bits2symbols_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int]),
        'gfdm_symbol_mapper_symbols2bits_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int]),
        'gfdm_modulator_work_chunks_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int]),
        'gfdm_transmitter_work_chunks_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int]),
        'gfdm_receiver_work_decide_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int]),
        'gfdm_resource_mapper_demap_chunks_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int]),
        'gfdm_receiver_work_strided_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_size_t, c_int, c_int]),
        'gfdm_burst_shaper_create': (c_int, [POINTER(c_void_p), c_int, c_int, c_float, c_float]),
        'gfdm_burst_shaper_destroy': (None, [c_void_p]),
        'gfdm_burst_shaper_pre_padding': (c_int, [c_void_p]),
        'gfdm_burst_shaper_post_padding': (c_int, [c_void_p]),
        'gfdm_burst_shaper_work_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int]),
        'gfdm_transmitter_work_shaped_batch': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int]),
        # sc16 sample format on the host side of a batch
        'gfdm_modulator_work_batch_sc16': (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int, c_int]),
        'gfdm_modulator_work_chunks_batch_sc16': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int]),
        'gfdm_receiver_work_batch_sc16': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int]),
        'gfdm_receiver_work_decide_batch_sc16': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int]),
    }

    def __init__(self, path):
        self.path = path
        self.dll = ctypes.CDLL(path)
        for name, (res, args) in self._PROTOS.items():
            fn = getattr(self.dll, name)  # AttributeError if a symbol is missing
            fn.restype = res
            fn.argtypes = args

    @property
    def exported_symbols(self):
        return sorted(self._PROTOS)

    def backend(self):
        return self.dll.gfdm_backend().decode()

    def device_count(self):
        return self.dll.gfdm_device_count()

    def set_device(self, dev):
        self.check(self.dll.gfdm_set_device(dev))

    def last_error(self):
        return self.dll.gfdm_last_error().decode(errors='replace')

    def check(self, status):
        if status == GFDM_OK:
            return
        msg = self.last_error()
        if status == GFDM_ERR_INVALID_ARGUMENT:
            raise ValueError(msg)  # pybind11: std::invalid_argument -> ValueError
        if status == GFDM_ERR_CUDA:
            raise GfdmCudaError(msg)
        raise RuntimeError(msg)    # pybind11: std::runtime_error -> RuntimeError

    def calculate_signal_energy(self, vec):
        v = _c64(vec)
        e = c_float()
        self.check(self.dll.gfdm_calculate_signal_energy(byref(e), _ptr(v), v.size))
        return e.value


_default = None


def load(path=None):
    """Load a library exporting the ABI.  Without a path: the product CUDA library.

    There is deliberately no fallback: if the CUDA library has not been built the
    import fails loudly instead of routing through some CPU path.
    """
    global _default
    if path is None:
        if _default is None:
            if not os.path.exists(DEFAULT_LIBRARY):
                raise ImportError('gfdm_b200: %s is missing -- build the CUDA library first '
                                  '(python -c "import __graft_entry__ as g; g.build()")' % DEFAULT_LIBRARY)
            _default = Library(DEFAULT_LIBRARY)
        return _default
    return Library(path)


class _Handle(object):
    _destroy = None

    def __init__(self, lib):
        self._lib = lib if lib is not None else load()
        self._dll = self._lib.dll
        self._h = c_void_p()

    def __del__(self):
        try:
            if getattr(self, '_h', None) is not None and self._h.value:
                getattr(self._dll, self._destroy)(self._h)
                self._h = c_void_p()
        except Exception:
            pass

    def _ck(self, st):
        self._lib.check(st)

    # -- stream / sync ------------------------------------------------------
    def set_stream(self, cuda_stream):
        self._ck(self._dll.gfdm_set_stream(self._h, c_void_p(int(cuda_stream))))

    def sync(self):
        self._ck(self._dll.gfdm_sync(self._h))

    def launch_count(self):
        return int(self._dll.gfdm_launch_count(self._h))

    def last_kernel(self):
        return self._dll.gfdm_last_kernel(self._h).decode()

    @staticmethod
    def _one_dim(arr, size, what, cls='Modulator'):
        if arr.ndim != 1:
            raise RuntimeError('Only ONE-dimensional vectors allowed!')
        if arr.size != size:
            raise RuntimeError('%s vector size(%d) MUST be equal to %s(%d)!' % (what, arr.size, cls, size))

    @staticmethod
    def _two_dim(arr, size):
        if arr.ndim != 2 or arr.shape[1] != size:
            raise RuntimeError('batch arrays MUST have shape [n_frames, %d], got %s' % (size, arr.shape))


class FFT(_Handle):
    _destroy = 'gfdm_fft_destroy'

    def __init__(self, fft_size, forward=True, lib=None):
        _Handle.__init__(self, lib)
        self.n = fft_size
        self._ck(self._dll.gfdm_fft_create(byref(self._h), fft_size, int(bool(forward))))

    def execute(self, x):
        x = _c64(x)
        shape = x.shape
        x = x.reshape(-1, self.n)
        out = np.empty_like(x)
        self._ck(self._dll.gfdm_fft_execute_batch(self._h, _ptr(out), _ptr(x), x.shape[0], MEM_HOST))
        return out.reshape(shape)

    def execute_ptr(self, out_ptr, in_ptr, n_transforms):
        self._ck(self._dll.gfdm_fft_execute_batch(self._h, c_void_p(out_ptr), c_void_p(in_ptr),
                                                  n_transforms, MEM_DEVICE))


class Modulator(_Handle):
    """modulator_kernel_cc / gfdm_python.Modulator (python/bindings/modulator_python.cc:34-59)."""
    _destroy = 'gfdm_modulator_destroy'

    def __init__(self, timeslots, subcarriers, overlap, taps, lib=None):
        _Handle.__init__(self, lib)
        t = _c64(taps)
        self._ck(self._dll.gfdm_modulator_create(byref(self._h), timeslots, subcarriers, overlap,
                                                 _ptr(t), t.size))
        self._ntaps = t.size

    def block_size(self):
        return self._dll.gfdm_modulator_block_size(self._h)

    def filter_taps(self):
        out = np.empty(self._ntaps, np.complex64)
        self._ck(self._dll.gfdm_modulator_filter_taps(self._h, _ptr(out)))
        return out

    def modulate(self, array):
        a = _c64(array)
        self._one_dim(a, self.block_size(), 'Input', 'Modulator.block_size')
        out = np.empty(a.size, np.complex64)
        self._ck(self._dll.gfdm_modulator_work(self._h, _ptr(out), _ptr(a)))
        return out

    def modulate_batch(self, array):
        a = _c64(array)
        self._two_dim(a, self.block_size())
        out = np.empty_like(a)
        self._ck(self._dll.gfdm_modulator_work_batch(self._h, _ptr(out), _ptr(a), a.shape[0], MEM_HOST))
        return out

    def modulate_batch_host_ptr(self, out_ptr, in_ptr, n_frames, mem=MEM_HOST):
        """HOST pointers; mem=MEM_HOST_ASYNC returns with the batch enqueued (pinned buffers; finish with sync())."""
        self._ck(self._dll.gfdm_modulator_work_batch(self._h, c_void_p(out_ptr), c_void_p(in_ptr),
                                                     n_frames, mem))

    def modulate_batch_sc16(self, array, scale):
        """[n_frames, block_size] symbols -> int16 I/Q samples [n_frames, block_size, 2] = round(x * scale), saturated."""
        a = _c64(array)
        self._two_dim(a, self.block_size())
        out = np.empty(a.shape + (2,), np.int16)
        self._ck(self._dll.gfdm_modulator_work_batch_sc16(self._h, _ptr(out), _ptr(a), scale, a.shape[0], MEM_HOST))
        return out

    def modulate_sc16_ptr(self, out_ptr, in_ptr, scale, n_frames, mem=MEM_DEVICE):
        self._ck(self._dll.gfdm_modulator_work_batch_sc16(self._h, c_void_p(out_ptr), c_void_p(in_ptr), scale, n_frames, mem))

    def modulate_chunks_batch_sc16(self, symbol_mapper, chunks, scale):
        c = _u8(chunks)
        self._two_dim(c, self.block_size())
        out = np.empty(c.shape + (2,), np.int16)
        self._ck(self._dll.gfdm_modulator_work_chunks_batch_sc16(self._h, symbol_mapper._h, _ptr(out), _ptr(c), scale,
                                                                 c.shape[0], MEM_HOST))
        return out

    def modulate_chunks_sc16_ptr(self, symbol_mapper, out_ptr, chunks_ptr, scale, n_frames, mem=MEM_DEVICE):
        self._ck(self._dll.gfdm_modulator_work_chunks_batch_sc16(self._h, symbol_mapper._h, c_void_p(out_ptr),
                                                                 c_void_p(chunks_ptr), scale, n_frames, mem))

    def modulate_ptr(self, out_ptr, in_ptr, n_frames):
        self._ck(self._dll.gfdm_modulator_work_batch(self._h, c_void_p(out_ptr), c_void_p(in_ptr),
                                                     n_frames, MEM_DEVICE))

    def modulate_chunks_batch(self, symbol_mapper, chunks):
        """chunks[n_frames, block_size] uint8 (constellation point indices) -> time samples."""
        c = _u8(chunks)
        self._two_dim(c, self.block_size())
        out = np.empty(c.shape, np.complex64)
        self._ck(self._dll.gfdm_modulator_work_chunks_batch(self._h, symbol_mapper._h, _ptr(out), _ptr(c),
                                                            c.shape[0], MEM_HOST))
        return out

    def modulate_chunks_batch_host_ptr(self, symbol_mapper, out_ptr, chunks_ptr, n_frames, mem=MEM_HOST):
        self._ck(self._dll.gfdm_modulator_work_chunks_batch(self._h, symbol_mapper._h, c_void_p(out_ptr),
                                                            c_void_p(chunks_ptr), n_frames, mem))

    def modulate_chunks_ptr(self, symbol_mapper, out_ptr, chunks_ptr, n_frames):
        self._ck(self._dll.gfdm_modulator_work_chunks_batch(self._h, symbol_mapper._h, c_void_p(out_ptr),
                                                            c_void_p(chunks_ptr), n_frames, MEM_DEVICE))


class Demodulator(_Handle):
    """receiver_kernel_cc / gfdm_python.Demodulator (python/bindings/demodulator_python.cc:35-205)."""
    _destroy = 'gfdm_receiver_destroy'

    def __init__(self, timeslots, subcarriers, overlap, taps, lib=None):
        _Handle.__init__(self, lib)
        t = _c64(taps)
        self._ck(self._dll.gfdm_receiver_create(byref(self._h), timeslots, subcarriers, overlap,
                                                _ptr(t), t.size))
        self._ntaps = t.size

    def timeslots(self):
        return self._dll.gfdm_receiver_timeslots(self._h)

    def subcarriers(self):
        return self._dll.gfdm_receiver_subcarriers(self._h)

    def overlap(self):
        return self._dll.gfdm_receiver_overlap(self._h)

    def block_size(self):
        return self._dll.gfdm_receiver_block_size(self._h)

    def filter_taps(self):
        out = np.empty(self._ntaps, np.complex64)
        self._ck(self._dll.gfdm_receiver_filter_taps(self._h, _ptr(out)))
        return out

    def ic_filter_taps(self):
        out = np.empty(self.timeslots(), np.complex64)
        self._ck(self._dll.gfdm_receiver_ic_filter_taps(self._h, _ptr(out)))
        return out

    def _unary(self, fn, array, cls='Modulator.block_size'):
        a = _c64(array)
        self._one_dim(a, self.block_size(), 'Input', cls)
        out = np.empty(a.size, np.complex64)
        self._ck(fn(self._h, _ptr(out), _ptr(a)))
        return out

    def _binary(self, fn, array, eq_arr):
        a, e = _c64(array), _c64(eq_arr)
        if a.ndim != 1 or e.ndim != 1:
            raise RuntimeError('Only ONE-dimensional vectors allowed!')
        self._one_dim(a, self.block_size(), 'Input', 'Demodulator.block_size')
        self._one_dim(e, self.block_size(), 'Channel', 'Demodulator.block_size')
        out = np.empty(a.size, np.complex64)
        self._ck(fn(self._h, _ptr(out), _ptr(a), _ptr(e)))
        return out

    def demodulate(self, array):
        return self._unary(self._dll.gfdm_receiver_work, array)

    def fft_filter_downsample(self, array):
        return self._unary(self._dll.gfdm_receiver_fft_filter_downsample, array)

    def transform_subcarriers_to_td(self, array):
        return self._unary(self._dll.gfdm_receiver_transform_subcarriers_to_td, array)

    def demodulate_equalize(self, array, eq_arr):
        return self._binary(self._dll.gfdm_receiver_work_equalize, array, eq_arr)

    def fft_equalize_filter_downsample(self, array, eq_arr):
        return self._binary(self._dll.gfdm_receiver_fft_equalize_filter_downsample, array, eq_arr)

    def cancel_sc_interference(self, array, eq_arr):
        return self._binary(self._dll.gfdm_receiver_cancel_sc_interference, array, eq_arr)

    # batched forms ----------------------------------------------------------
    def demodulate_batch(self, array, eq_arr=None):
        a = _c64(array)
        self._two_dim(a, self.block_size())
        e = None
        if eq_arr is not None:
            e = _c64(eq_arr)
            self._two_dim(e, self.block_size())
        out = np.empty_like(a)
        self._ck(self._dll.gfdm_receiver_work_batch(self._h, _ptr(out), _ptr(a),
                                                    _ptr(e) if e is not None else None,
                                                    a.shape[0], MEM_HOST))
        return out

    def fft_filter_downsample_batch(self, array, eq_arr=None):
        a = _c64(array)
        self._two_dim(a, self.block_size())
        e = None
        if eq_arr is not None:
            e = _c64(eq_arr)
            self._two_dim(e, self.block_size())
        out = np.empty_like(a)
        self._ck(self._dll.gfdm_receiver_fft_filter_downsample_batch(
            self._h, _ptr(out), _ptr(a), _ptr(e) if e is not None else None, a.shape[0], MEM_HOST))
        return out

    def transform_subcarriers_to_td_batch(self, array):
        a = _c64(array)
        self._two_dim(a, self.block_size())
        out = np.empty_like(a)
        self._ck(self._dll.gfdm_receiver_transform_subcarriers_to_td_batch(
            self._h, _ptr(out), _ptr(a), a.shape[0], MEM_HOST))
        return out

    def cancel_sc_interference_batch(self, td, fd):
        a, e = _c64(td), _c64(fd)
        self._two_dim(a, self.block_size())
        self._two_dim(e, self.block_size())
        out = np.empty_like(a)
        self._ck(self._dll.gfdm_receiver_cancel_sc_interference_batch(
            self._h, _ptr(out), _ptr(a), _ptr(e), a.shape[0], MEM_HOST))
        return out

    def demodulate_batch_host_ptr(self, out_ptr, in_ptr, eq_ptr, n_frames, mem=MEM_HOST):
        """HOST pointers; mem=MEM_HOST_ASYNC returns with the batch enqueued (pinned buffers; finish with sync())."""
        self._ck(self._dll.gfdm_receiver_work_batch(self._h, c_void_p(out_ptr), c_void_p(in_ptr),
                                                    c_void_p(eq_ptr) if eq_ptr else None,
                                                    n_frames, mem))

    def demodulate_strided_batch(self, frames, offset, eq_arr=None):
        """frames[n, stride] still carrying prefix / suffix: demodulates frames[:, offset:offset + block_size] in place."""
        a = _c64(frames)
        if a.ndim != 2 or a.shape[1] < offset + self.block_size():
            raise RuntimeError('frames MUST have shape [n_frames, >= offset + block_size]')
        eq = None
        if eq_arr is not None:
            eq = _c64(eq_arr)
            self._two_dim(eq, self.block_size())
        out = np.empty((a.shape[0], self.block_size()), np.complex64)
        self._ck(self._dll.gfdm_receiver_work_strided_batch(self._h, _ptr(out), _ptr(a), _ptr(eq) if eq is not None else None,
                                                            a.shape[1], offset, a.shape[0], MEM_HOST))
        return out

    def demodulate_strided_ptr(self, out_ptr, in_ptr, eq_ptr, stride, offset, n_frames, mem=MEM_DEVICE):
        self._ck(self._dll.gfdm_receiver_work_strided_batch(self._h, c_void_p(out_ptr), c_void_p(in_ptr),
                                                            c_void_p(eq_ptr) if eq_ptr else None, stride, offset, n_frames, mem))

    @staticmethod
    def _iq(array, size):
        a = np.ascontiguousarray(array, dtype=np.int16)
        if a.ndim != 3 or a.shape[1] != size or a.shape[2] != 2:
            raise RuntimeError('sc16 batches MUST have shape [n_frames, %d, 2], got %s' % (size, a.shape))
        return a

    def demodulate_batch_sc16(self, iq, scale, eq_arr=None):
        """int16 I/Q samples [n_frames, block_size, 2] (value = sample * scale) -> soft symbols."""
        a = self._iq(iq, self.block_size())
        eq = None
        if eq_arr is not None:
            eq = _c64(eq_arr)
            self._two_dim(eq, self.block_size())
        out = np.empty(a.shape[:2], np.complex64)
        self._ck(self._dll.gfdm_receiver_work_batch_sc16(self._h, _ptr(out), _ptr(a), _ptr(eq) if eq is not None else None,
                                                         scale, a.shape[0], MEM_HOST))
        return out

    def demodulate_sc16_ptr(self, out_ptr, in_ptr, eq_ptr, scale, n_frames, mem=MEM_DEVICE):
        self._ck(self._dll.gfdm_receiver_work_batch_sc16(self._h, c_void_p(out_ptr), c_void_p(in_ptr),
                                                         c_void_p(eq_ptr) if eq_ptr else None, scale, n_frames, mem))

    def demodulate_decide_batch_sc16(self, symbol_mapper, iq, scale, eq_arr=None):
        a = self._iq(iq, self.block_size())
        eq = None
        if eq_arr is not None:
            eq = _c64(eq_arr)
            self._two_dim(eq, self.block_size())
        out = np.empty(a.shape[:2], np.uint8)
        self._ck(self._dll.gfdm_receiver_work_decide_batch_sc16(self._h, symbol_mapper._h, _ptr(out), _ptr(a),
                                                                _ptr(eq) if eq is not None else None, scale, a.shape[0],
                                                                MEM_HOST))
        return out

    def demodulate_decide_sc16_ptr(self, symbol_mapper, out_ptr, in_ptr, eq_ptr, scale, n_frames, mem=MEM_DEVICE):
        self._ck(self._dll.gfdm_receiver_work_decide_batch_sc16(self._h, symbol_mapper._h, c_void_p(out_ptr), c_void_p(in_ptr),
                                                                c_void_p(eq_ptr) if eq_ptr else None, scale, n_frames, mem))

    def demodulate_decide_batch(self, symbol_mapper, array, eq_arr=None):
        """time samples [n_frames, block_size] -> hard decisions (uint8 point indices) on the full grid."""
        a = _c64(array)
        self._two_dim(a, self.block_size())
        eq = None
        if eq_arr is not None:
            eq = _c64(eq_arr)
            self._two_dim(eq, self.block_size())
        out = np.empty(a.shape, np.uint8)
        self._ck(self._dll.gfdm_receiver_work_decide_batch(self._h, symbol_mapper._h, _ptr(out), _ptr(a),
                                                           _ptr(eq) if eq is not None else None,
                                                           a.shape[0], MEM_HOST))
        return out

    def demodulate_decide_batch_host_ptr(self, symbol_mapper, out_ptr, in_ptr, eq_ptr, n_frames, mem=MEM_HOST):
        self._ck(self._dll.gfdm_receiver_work_decide_batch(self._h, symbol_mapper._h, c_void_p(out_ptr),
                                                           c_void_p(in_ptr), c_void_p(eq_ptr) if eq_ptr else None,
                                                           n_frames, mem))

    def demodulate_decide_ptr(self, symbol_mapper, out_ptr, in_ptr, eq_ptr, n_frames):
        self._ck(self._dll.gfdm_receiver_work_decide_batch(self._h, symbol_mapper._h, c_void_p(out_ptr),
                                                           c_void_p(in_ptr), c_void_p(eq_ptr) if eq_ptr else None,
                                                           n_frames, MEM_DEVICE))

    def demodulate_ptr(self, out_ptr, in_ptr, eq_ptr, n_frames, mem=MEM_DEVICE):
        self._ck(self._dll.gfdm_receiver_work_batch(self._h, c_void_p(out_ptr), c_void_p(in_ptr),
                                                    c_void_p(eq_ptr) if eq_ptr else None,
                                                    n_frames, mem))


def qpsk_constellation():
    """points()/decision rule of gr::digital::constellation_qpsk (GNU Radio 3.9)."""
    s = np.float32(0.707107)
    pts = np.array([-s - 1j * s, s - 1j * s, -s + 1j * s, s + 1j * s], dtype=np.complex64)
    return pts, DECISION_QPSK_SIGN


class Advanced_receiver(_Handle):
    """advanced_receiver_kernel_cc (include/gfdm/advanced_receiver_kernel_cc.h:37-61).

    ``constellation`` is ``(points, decision_rule)``; default: GNU Radio's QPSK.
    """
    _destroy = 'gfdm_advanced_receiver_destroy'

    def __init__(self, timeslots, subcarriers, overlap, taps, subcarrier_map, ic_iter,
                 constellation=None, do_phase_compensation=0, lib=None):
        _Handle.__init__(self, lib)
        t = _c64(taps)
        smap = np.ascontiguousarray(subcarrier_map, dtype=np.int32)
        pts, rule = constellation if constellation is not None else qpsk_constellation()
        pts = _c64(pts)
        c = _Constellation(pts.ctypes.data, pts.size, int(rule))
        self._ck(self._dll.gfdm_advanced_receiver_create(
            byref(self._h), timeslots, subcarriers, overlap, _ptr(t), t.size, _ptr(smap), smap.size,
            ic_iter, byref(c), do_phase_compensation))

    def block_size(self):
        return self._dll.gfdm_advanced_receiver_block_size(self._h)

    def set_ic(self, ic_iter):
        self._ck(self._dll.gfdm_advanced_receiver_set_ic(self._h, ic_iter))

    def get_ic(self):
        return self._dll.gfdm_advanced_receiver_get_ic(self._h)

    def set_phase_compensation(self, v):
        self._ck(self._dll.gfdm_advanced_receiver_set_phase_compensation(self._h, v))

    def get_phase_compensation(self):
        return self._dll.gfdm_advanced_receiver_get_phase_compensation(self._h)

    def demodulate(self, array):
        a = _c64(array)
        self._one_dim(a, self.block_size(), 'Input', 'Advanced_receiver.block_size')
        out = np.empty(a.size, np.complex64)
        self._ck(self._dll.gfdm_advanced_receiver_work(self._h, _ptr(out), _ptr(a)))
        return out

    def demodulate_equalize(self, array, eq_arr):
        a, e = _c64(array), _c64(eq_arr)
        self._one_dim(a, self.block_size(), 'Input', 'Advanced_receiver.block_size')
        self._one_dim(e, self.block_size(), 'Channel', 'Advanced_receiver.block_size')
        out = np.empty(a.size, np.complex64)
        self._ck(self._dll.gfdm_advanced_receiver_work_equalize(self._h, _ptr(out), _ptr(a), _ptr(e)))
        return out

    def demodulate_batch(self, array, eq_arr=None):
        a = _c64(array)
        self._two_dim(a, self.block_size())
        e = None
        if eq_arr is not None:
            e = _c64(eq_arr)
            self._two_dim(e, self.block_size())
        out = np.empty_like(a)
        self._ck(self._dll.gfdm_advanced_receiver_work_batch(
            self._h, _ptr(out), _ptr(a), _ptr(e) if e is not None else None, a.shape[0], MEM_HOST))
        return out

    def demodulate_ptr(self, out_ptr, in_ptr, eq_ptr, n_frames, mem=MEM_DEVICE):
        self._ck(self._dll.gfdm_advanced_receiver_work_batch(
            self._h, c_void_p(out_ptr), c_void_p(in_ptr), c_void_p(eq_ptr) if eq_ptr else None,
            n_frames, mem))


class Resource_mapper(_Handle):
    """resource_mapper_kernel_cc / gfdm_python.Resource_mapper (resource_mapper_python.cc:34-85)."""
    _destroy = 'gfdm_resource_mapper_destroy'

    def __init__(self, timeslots, subcarriers, active_subcarriers, subcarrier_map, per_timeslot=True,
                 is_mapper=True, lib=None):
        _Handle.__init__(self, lib)
        smap = np.ascontiguousarray(subcarrier_map, dtype=np.int32)
        self._ck(self._dll.gfdm_resource_mapper_create(byref(self._h), timeslots, subcarriers,
                                                       active_subcarriers, _ptr(smap), smap.size,
                                                       int(bool(per_timeslot)), int(bool(is_mapper))))

    def block_size(self):
        return int(self._dll.gfdm_resource_mapper_block_size(self._h))

    def frame_size(self):
        return int(self._dll.gfdm_resource_mapper_frame_size(self._h))

    def input_vector_size(self):
        return int(self._dll.gfdm_resource_mapper_input_vector_size(self._h))

    def output_vector_size(self):
        return int(self._dll.gfdm_resource_mapper_output_vector_size(self._h))

    def map_to_resources(self, array):
        a = _c64(array)
        self._one_dim(a, self.block_size(), 'Input', 'Modulator.block_size')
        out = np.empty(self.frame_size(), np.complex64)
        self._ck(self._dll.gfdm_resource_mapper_map_to_resources(self._h, _ptr(out), _ptr(a),
                                                                 self.block_size()))
        return out

    def demap_from_resources(self, array):
        a = _c64(array)
        self._one_dim(a, self.frame_size(), 'Input', 'Modulator.block_size')
        out = np.empty(self.block_size(), np.complex64)
        self._ck(self._dll.gfdm_resource_mapper_demap_from_resources(self._h, _ptr(out), _ptr(a),
                                                                     self.block_size()))
        return out

    # the kernel-level calls with an explicit symbol count (lib/resource_mapper_kernel_cc.cc:74-106)
    def map_to_resources_n(self, array, ninput_size):
        a = _c64(array)
        out = np.empty(self.frame_size(), np.complex64)
        self._ck(self._dll.gfdm_resource_mapper_map_to_resources(self._h, _ptr(out), _ptr(a), ninput_size))
        return out

    def demap_from_resources_n(self, array, noutput_size, pad=1):
        a = _c64(array)
        out = np.zeros(noutput_size + pad, np.complex64)
        self._ck(self._dll.gfdm_resource_mapper_demap_from_resources(self._h, _ptr(out), _ptr(a), noutput_size))
        return out

    def map_to_resources_batch(self, array):
        a = _c64(array)
        if a.ndim != 2 or a.shape[1] > self.block_size():
            raise RuntimeError('batch arrays MUST have shape [n_frames, <= %d]' % self.block_size())
        out = np.empty((a.shape[0], self.frame_size()), np.complex64)
        self._ck(self._dll.gfdm_resource_mapper_map_to_resources_batch(
            self._h, _ptr(out), _ptr(a), a.shape[1], a.shape[0], MEM_HOST))
        return out

    def demap_from_resources_batch(self, array, size_per_frame=None):
        a = _c64(array)
        self._two_dim(a, self.frame_size())
        n = self.block_size() if size_per_frame is None else size_per_frame
        out = np.empty((a.shape[0], n), np.complex64)
        self._ck(self._dll.gfdm_resource_mapper_demap_from_resources_batch(
            self._h, _ptr(out), _ptr(a), n, a.shape[0], MEM_HOST))
        return out

    def demap_chunks_batch(self, chunks, size_per_frame=None):
        """demap_from_resources on a grid of uint8 chunks [n_frames, frame_size]."""
        c = _u8(chunks)
        self._two_dim(c, self.frame_size())
        n = self.block_size() if size_per_frame is None else size_per_frame
        out = np.empty((c.shape[0], n), np.uint8)
        self._ck(self._dll.gfdm_resource_mapper_demap_chunks_batch(self._h, _ptr(out), _ptr(c), n, c.shape[0],
                                                                   MEM_HOST))
        return out

    def demap_chunks_ptr(self, out_ptr, in_ptr, size_per_frame, n_frames):
        self._ck(self._dll.gfdm_resource_mapper_demap_chunks_batch(
            self._h, c_void_p(out_ptr), c_void_p(in_ptr), size_per_frame, n_frames, MEM_DEVICE))

    def map_ptr(self, out_ptr, in_ptr, size_per_frame, n_frames):
        self._ck(self._dll.gfdm_resource_mapper_map_to_resources_batch(
            self._h, c_void_p(out_ptr), c_void_p(in_ptr), size_per_frame, n_frames, MEM_DEVICE))

    def demap_ptr(self, out_ptr, in_ptr, size_per_frame, n_frames):
        self._ck(self._dll.gfdm_resource_mapper_demap_from_resources_batch(
            self._h, c_void_p(out_ptr), c_void_p(in_ptr), size_per_frame, n_frames, MEM_DEVICE))


class Cyclic_prefixer(_Handle):
    """add_cyclic_prefix_cc / gfdm_python.Cyclic_prefixer (cyclic_prefix_python.cc:34-94)."""
    _destroy = 'gfdm_cyclic_prefixer_destroy'

    def __init__(self, block_len, cp_len, cs_len, ramp_len, window_taps, cyclic_shift=0, lib=None):
        _Handle.__init__(self, lib)
        w = _c64(window_taps)
        self._ck(self._dll.gfdm_cyclic_prefixer_create(byref(self._h), block_len, cp_len, cs_len,
                                                       ramp_len, _ptr(w), w.size, cyclic_shift))

    def block_size(self):
        return self._dll.gfdm_cyclic_prefixer_block_size(self._h)

    def frame_size(self):
        return self._dll.gfdm_cyclic_prefixer_frame_size(self._h)

    def cyclic_shift(self):
        return self._dll.gfdm_cyclic_prefixer_cyclic_shift(self._h)

    def add_cyclic_prefix(self, array):
        a = _c64(array)
        self._one_dim(a, self.block_size(), 'Input', 'Cyclic_prefix.block_size')
        out = np.empty(self.frame_size(), np.complex64)
        self._ck(self._dll.gfdm_cyclic_prefixer_work(self._h, _ptr(out), _ptr(a)))
        return out

    def add_cyclic_prefix_shifted(self, array, cyclic_shift):
        a = _c64(array)
        self._one_dim(a, self.block_size(), 'Input', 'Cyclic_prefix.block_size')
        out = np.empty(self.frame_size(), np.complex64)
        self._ck(self._dll.gfdm_cyclic_prefixer_add_cyclic_prefix(self._h, _ptr(out), _ptr(a), cyclic_shift))
        return out

    def remove_cyclic_prefix(self, array):
        a = _c64(array)
        if a.ndim != 1:
            raise RuntimeError('Only ONE-dimensional vectors allowed!')
        if a.size != self.frame_size():
            raise RuntimeError('Input vector size(%d) MUST be equal to Cyclic_prefix.frame_size(%d)!'
                               % (a.size, self.block_size()))
        out = np.empty(self.block_size(), np.complex64)
        self._ck(self._dll.gfdm_cyclic_prefixer_remove_cyclic_prefix(self._h, _ptr(out), _ptr(a)))
        return out

    def add_cyclic_prefix_batch(self, array, cyclic_shift=None):
        a = _c64(array)
        self._two_dim(a, self.block_size())
        out = np.empty((a.shape[0], self.frame_size()), np.complex64)
        s = self.cyclic_shift() if cyclic_shift is None else cyclic_shift
        self._ck(self._dll.gfdm_cyclic_prefixer_add_cyclic_prefix_batch(
            self._h, _ptr(out), _ptr(a), s, a.shape[0], MEM_HOST))
        return out

    def remove_cyclic_prefix_batch(self, array):
        a = _c64(array)
        self._two_dim(a, self.frame_size())
        out = np.empty((a.shape[0], self.block_size()), np.complex64)
        self._ck(self._dll.gfdm_cyclic_prefixer_remove_cyclic_prefix_batch(
            self._h, _ptr(out), _ptr(a), a.shape[0], MEM_HOST))
        return out

    def add_ptr(self, out_ptr, in_ptr, cyclic_shift, n_frames):
        self._ck(self._dll.gfdm_cyclic_prefixer_add_cyclic_prefix_batch(
            self._h, c_void_p(out_ptr), c_void_p(in_ptr), cyclic_shift, n_frames, MEM_DEVICE))

    def remove_ptr(self, out_ptr, in_ptr, n_frames):
        self._ck(self._dll.gfdm_cyclic_prefixer_remove_cyclic_prefix_batch(
            self._h, c_void_p(out_ptr), c_void_p(in_ptr), n_frames, MEM_DEVICE))


class Preamble_channel_estimator(_Handle):
    """preamble_channel_estimator_cc (preamble_channel_estimator_python.cc:34-99)."""
    _destroy = 'gfdm_channel_estimator_destroy'

    def __init__(self, timeslots, subcarriers, active_subcarriers, is_dc_free, which_estimator,
                 preamble, lib=None):
        _Handle.__init__(self, lib)
        p = _c64(preamble)
        self._ck(self._dll.gfdm_channel_estimator_create(byref(self._h), timeslots, subcarriers,
                                                         active_subcarriers, int(bool(is_dc_free)),
                                                         which_estimator, _ptr(p), p.size))

    def timeslots(self):
        return self._dll.gfdm_channel_estimator_timeslots(self._h)

    def subcarriers(self):
        return self._dll.gfdm_channel_estimator_fft_len(self._h)

    def active_subcarriers(self):
        return self._dll.gfdm_channel_estimator_active_subcarriers(self._h)

    def frame_len(self):
        return self._dll.gfdm_channel_estimator_frame_len(self._h)

    def is_dc_free(self):
        return bool(self._dll.gfdm_channel_estimator_is_dc_free(self._h))

    def preamble_filter_taps(self):
        out = np.empty(9, np.float32)
        self._ck(self._dll.gfdm_channel_estimator_preamble_filter_taps(self._h, _ptr(out)))
        return out

    def _check_in(self, a):
        if a.ndim != 1:
            raise RuntimeError('Only ONE-dimensional vectors allowed!')
        if a.size != 2 * self.subcarriers():
            raise RuntimeError('Input vector size(%d) MUST be equal to 2 * subcarriers(%d)!'
                               % (a.size, 2 * self.subcarriers()))

    def estimate_frame(self, array, fill=0.0):
        a = _c64(array)
        self._check_in(a)
        out = np.full(self.frame_len(), fill, np.complex64)
        self._ck(self._dll.gfdm_channel_estimator_estimate_frame(self._h, _ptr(out), _ptr(a)))
        return out

    def estimate_snr(self, array):
        a = _c64(array)
        self._check_in(a)
        snr = c_float()
        self._ck(self._dll.gfdm_channel_estimator_estimate_snr(self._h, byref(snr), None, _ptr(a)))
        return snr.value

    def estimate_snr_cnrs(self, array):
        a = _c64(array)
        self._check_in(a)
        snr = c_float()
        cnrs = np.empty(self.active_subcarriers(), np.float32)
        self._ck(self._dll.gfdm_channel_estimator_estimate_snr(self._h, byref(snr), _ptr(cnrs), _ptr(a)))
        return snr.value, cnrs

    def estimate_preamble_channel(self, array):
        a = _c64(array)
        self._check_in(a)
        out = np.empty(self.subcarriers(), np.complex64)
        self._ck(self._dll.gfdm_channel_estimator_estimate_preamble_channel(self._h, _ptr(out), _ptr(a)))
        return out

    def filter_preamble_estimate(self, estimate):
        a = _c64(estimate)
        n = self.active_subcarriers() + (1 if self.is_dc_free() else 0)
        out = np.empty(n, np.complex64)
        self._ck(self._dll.gfdm_channel_estimator_filter_preamble_estimate(self._h, _ptr(out), _ptr(a)))
        return out

    def interpolate_frame(self, estimate, fill=0.0):
        a = _c64(estimate)
        out = np.full(self.frame_len(), fill, np.complex64)
        self._ck(self._dll.gfdm_channel_estimator_interpolate_frame(self._h, _ptr(out), _ptr(a)))
        return out

    def prepare_for_zf(self, frame_estimate):
        a = _c64(frame_estimate)
        out = np.empty(self.frame_len(), np.complex64)
        self._ck(self._dll.gfdm_channel_estimator_prepare_for_zf(self._h, _ptr(out), _ptr(a)))
        return out

    def estimate_frame_batch(self, array, fill=0.0):
        a = _c64(array)
        self._two_dim(a, 2 * self.subcarriers())
        out = np.full((a.shape[0], self.frame_len()), fill, np.complex64)
        self._ck(self._dll.gfdm_channel_estimator_estimate_frame_batch(
            self._h, _ptr(out), _ptr(a), a.shape[0], MEM_HOST))
        return out

    def estimate_snr_batch(self, array):
        a = _c64(array)
        self._two_dim(a, 2 * self.subcarriers())
        snr = np.empty(a.shape[0], np.float32)
        cnrs = np.empty((a.shape[0], self.active_subcarriers()), np.float32)
        self._ck(self._dll.gfdm_channel_estimator_estimate_snr_batch(
            self._h, _ptr(snr), _ptr(cnrs), _ptr(a), a.shape[0], MEM_HOST))
        return snr, cnrs

    def estimate_frame_ptr(self, out_ptr, in_ptr, n_frames, mem=MEM_DEVICE):
        self._ck(self._dll.gfdm_channel_estimator_estimate_frame_batch(
            self._h, c_void_p(out_ptr), c_void_p(in_ptr), n_frames, mem))


class Transmitter(_Handle):
    """transmitter_kernel (include/gfdm/transmitter_kernel.h:43-69)."""
    _destroy = 'gfdm_transmitter_destroy'

    def __init__(self, timeslots, subcarriers, active_subcarriers, cp_len, cs_len, ramp_len,
                 subcarrier_map, per_timeslot, overlap, frequency_taps, window_taps, cyclic_shifts,
                 preambles, lib=None):
        _Handle.__init__(self, lib)
        smap = np.ascontiguousarray(subcarrier_map, dtype=np.int32)
        t = _c64(frequency_taps)
        w = _c64(window_taps)
        shifts = np.ascontiguousarray(cyclic_shifts, dtype=np.int32)
        pre = [_c64(p) for p in preambles]
        ptrs = (c_void_p * max(1, len(pre)))(*[p.ctypes.data for p in pre])
        sizes = np.array([p.size for p in pre], dtype=np.int32)
        self._ck(self._dll.gfdm_transmitter_create(
            byref(self._h), timeslots, subcarriers, active_subcarriers, cp_len, cs_len, ramp_len,
            _ptr(smap), smap.size, int(bool(per_timeslot)), overlap, _ptr(t), t.size, _ptr(w), w.size,
            _ptr(shifts), shifts.size, ptrs, _ptr(sizes), len(pre)))
        self._block = timeslots * subcarriers

    def set_chain_fusion(self, on):
        """CUDA library: run the chain as one kernel (default) or as four separate kernels."""
        self._ck(self._dll.gfdm_transmitter_set_chain_fusion(self._h, int(bool(on))))

    def input_vector_size(self):
        return self._dll.gfdm_transmitter_input_vector_size(self._h)

    def output_vector_size(self):
        return self._dll.gfdm_transmitter_output_vector_size(self._h)

    def cyclic_shifts(self):
        n = self._dll.gfdm_transmitter_n_cyclic_shifts(self._h)
        out = np.empty(n, np.int32)
        self._ck(self._dll.gfdm_transmitter_cyclic_shifts(self._h, _ptr(out)))
        return [int(v) for v in out]

    def generic_work(self, array, ninput_size=None):
        a = _c64(array)
        n = a.size if ninput_size is None else ninput_size
        out = np.empty(self.output_vector_size(), np.complex64)
        self._ck(self._dll.gfdm_transmitter_work(self._h, _ptr(out), _ptr(a), n))
        return out

    def modulate(self, array, ninput_size=None):
        a = _c64(array)
        n = a.size if ninput_size is None else ninput_size
        out = np.empty(self._block, np.complex64)
        self._ck(self._dll.gfdm_transmitter_modulate(self._h, _ptr(out), _ptr(a), n))
        return out

    def add_frame(self, array, cyclic_shift):
        a = _c64(array)
        self._one_dim(a, self._block, 'Input', 'Transmitter.block_size')
        out = np.empty(self.output_vector_size(), np.complex64)
        self._ck(self._dll.gfdm_transmitter_add_frame(self._h, _ptr(out), _ptr(a), cyclic_shift))
        return out

    def work_batch(self, array):
        a = _c64(array)
        if a.ndim != 2:
            raise RuntimeError('batch arrays MUST be two-dimensional')
        out = np.empty((a.shape[0], self.output_vector_size()), np.complex64)
        self._ck(self._dll.gfdm_transmitter_work_batch(self._h, _ptr(out), _ptr(a), a.shape[1],
                                                       a.shape[0], MEM_HOST))
        return out

    def work_all_batch(self, array):
        a = _c64(array)
        if a.ndim != 2:
            raise RuntimeError('batch arrays MUST be two-dimensional')
        n_sh = self._dll.gfdm_transmitter_n_cyclic_shifts(self._h)
        out = np.empty((n_sh, a.shape[0], self.output_vector_size()), np.complex64)
        self._ck(self._dll.gfdm_transmitter_work_all_batch(self._h, _ptr(out), _ptr(a), a.shape[1],
                                                           a.shape[0], MEM_HOST))
        return out

    def work_chunks_batch(self, symbol_mapper, chunks):
        """chunks[n_frames, ninput_size] uint8 -> frames of cyclic_shifts[0]."""
        c = _u8(chunks)
        if c.ndim != 2:
            raise RuntimeError('batch arrays MUST be two-dimensional')
        out = np.empty((c.shape[0], self.output_vector_size()), np.complex64)
        self._ck(self._dll.gfdm_transmitter_work_chunks_batch(self._h, symbol_mapper._h, _ptr(out), _ptr(c),
                                                              c.shape[1], c.shape[0], MEM_HOST))
        return out

    def work_shaped_batch(self, shaper, array, all_antennas=False):
        """transmitter chain + short_burst_shaper epilogue: [n_frames, pre + output_vector_size + post] (per antenna)."""
        a = _c64(array)
        if a.ndim != 2:
            raise RuntimeError('batch arrays MUST have shape [n_frames, ninput_size]')
        n_ant = len(self.cyclic_shifts()) if all_antennas else 1
        row = shaper.pre_padding + self.output_vector_size() + shaper.post_padding
        out = np.empty((n_ant, a.shape[0], row), np.complex64)
        self._ck(self._dll.gfdm_transmitter_work_shaped_batch(self._h, shaper._h, _ptr(out), _ptr(a), a.shape[1], a.shape[0],
                                                              int(bool(all_antennas)), MEM_HOST))
        return out if all_antennas else out[0]

    def work_shaped_ptr(self, shaper, out_ptr, in_ptr, ninput_size, n_frames, all_antennas=False, mem=MEM_DEVICE):
        self._ck(self._dll.gfdm_transmitter_work_shaped_batch(self._h, shaper._h, c_void_p(out_ptr), c_void_p(in_ptr), ninput_size,
                                                              n_frames, int(bool(all_antennas)), mem))

    def work_chunks_ptr(self, symbol_mapper, out_ptr, chunks_ptr, ninput_size, n_frames):
        self._ck(self._dll.gfdm_transmitter_work_chunks_batch(self._h, symbol_mapper._h, c_void_p(out_ptr),
                                                              c_void_p(chunks_ptr), ninput_size, n_frames, MEM_DEVICE))

    def work_ptr(self, out_ptr, in_ptr, ninput_size, n_frames, all_antennas=False, mem=MEM_DEVICE):
        fn = self._dll.gfdm_transmitter_work_all_batch if all_antennas else self._dll.gfdm_transmitter_work_batch
        self._ck(fn(self._h, c_void_p(out_ptr), c_void_p(in_ptr), ninput_size, n_frames, mem))


class Burst_shaper(_Handle):
    """short_burst_shaper's sample path (lib/short_burst_shaper_impl.cc:161-182): [pre zeros | in * scale | post zeros]."""
    _destroy = 'gfdm_burst_shaper_destroy'

    def __init__(self, pre_padding, post_padding, scale=1.0, lib=None):
        _Handle.__init__(self, lib)
        sc = complex(scale)
        self.pre_padding, self.post_padding, self.scale = pre_padding, post_padding, np.complex64(sc)
        self._ck(self._dll.gfdm_burst_shaper_create(byref(self._h), pre_padding, post_padding, sc.real, sc.imag))

    def work_batch(self, array):
        a = _c64(array)
        if a.ndim != 2:
            raise RuntimeError('batch arrays MUST have shape [n_bursts, burst_len]')
        out = np.empty((a.shape[0], self.pre_padding + a.shape[1] + self.post_padding), np.complex64)
        self._ck(self._dll.gfdm_burst_shaper_work_batch(self._h, _ptr(out), _ptr(a), a.shape[1], a.shape[0], MEM_HOST))
        return out

    def work_ptr(self, out_ptr, in_ptr, burst_len, n_bursts, mem=MEM_DEVICE):
        self._ck(self._dll.gfdm_burst_shaper_work_batch(self._h, c_void_p(out_ptr), c_void_p(in_ptr), burst_len, n_bursts, mem))


class Remove_prefix(_Handle):
    """remove_prefix_cc (lib/remove_prefix_cc_impl.cc:84-115) as a batched gather."""
    _destroy = 'gfdm_remove_prefix_destroy'

    def __init__(self, frame_len, block_len, offset, lib=None):
        _Handle.__init__(self, lib)
        self.frame_len, self.block_len, self.offset = frame_len, block_len, offset
        self._ck(self._dll.gfdm_remove_prefix_create(byref(self._h), frame_len, block_len, offset))

    def work_batch(self, array):
        a = _c64(array)
        self._two_dim(a, self.frame_len)
        out = np.empty((a.shape[0], self.block_len), np.complex64)
        self._ck(self._dll.gfdm_remove_prefix_work_batch(self._h, _ptr(out), _ptr(a), a.shape[0], MEM_HOST))
        return out

    def work_ptr(self, out_ptr, in_ptr, n_frames, mem=MEM_DEVICE):
        self._ck(self._dll.gfdm_remove_prefix_work_batch(self._h, c_void_p(out_ptr), c_void_p(in_ptr), n_frames, mem))


class Extract_burst(_Handle):
    """extract_burst_cc (lib/extract_burst_cc_impl.cc:117-242): one general_work() call per ``work``."""
    _destroy = 'gfdm_extract_burst_destroy'

    def __init__(self, burst_len, tag_backoff, activate_cfo_correction=False, lib=None):
        _Handle.__init__(self, lib)
        self.burst_len = burst_len
        self._ck(self._dll.gfdm_extract_burst_create(byref(self._h), burst_len, tag_backoff,
                                                     int(bool(activate_cfo_correction))))

    def activate_cfo_compensation(self, on):
        self._ck(self._dll.gfdm_extract_burst_activate_cfo_compensation(self._h, int(bool(on))))

    def _tags(self, burst_starts, scale_factors, phase_rotations):
        st = np.ascontiguousarray(burst_starts, dtype=np.int64)
        sc = None if scale_factors is None else np.ascontiguousarray(scale_factors, dtype=np.float32)
        pr = None if phase_rotations is None else _c64(phase_rotations)
        for x in (sc, pr):
            if x is not None and x.size != st.size:
                raise RuntimeError('tag arrays MUST have the same length')
        return st, sc, pr

    def work(self, stream, burst_starts, scale_factors=None, phase_rotations=None, max_bursts=None):
        """Returns (bursts[n_produced, burst_len], n_consumed)."""
        x = _c64(stream).ravel()
        st, sc, pr = self._tags(burst_starts, scale_factors, phase_rotations)
        mb = st.size if max_bursts is None else max_bursts
        out = np.empty((max(mb, 1), self.burst_len), np.complex64)
        n_prod, n_cons = c_int(), c_longlong()
        self._ck(self._dll.gfdm_extract_burst_work(
            self._h, _ptr(out), mb, _ptr(x), x.size, _ptr(st), _ptr(sc) if sc is not None else None,
            _ptr(pr) if pr is not None else None, st.size, byref(n_prod), byref(n_cons), MEM_HOST))
        return out[:n_prod.value].copy(), n_cons.value

    def work_ptr(self, out_ptr, max_bursts, in_ptr, n_in, burst_starts, scale_factors=None, phase_rotations=None):
        st, sc, pr = self._tags(burst_starts, scale_factors, phase_rotations)
        n_prod, n_cons = c_int(), c_longlong()
        self._ck(self._dll.gfdm_extract_burst_work(
            self._h, c_void_p(out_ptr), max_bursts, c_void_p(in_ptr), n_in, _ptr(st),
            _ptr(sc) if sc is not None else None, _ptr(pr) if pr is not None else None, st.size,
            byref(n_prod), byref(n_cons), MEM_DEVICE))
        return n_prod.value, n_cons.value


class Symbol_mapper(_Handle):
    """Symbol mapping either side of the path: pygfdm bits2symbols / symbols2bits
    (python/pygfdm/symbolmapping.py:27-47) and the chunk <-> point mapping of gr-digital's
    chunks_to_symbols / constellation decoder.  ``constellation`` is ``(points, decision_rule)``."""
    _destroy = 'gfdm_symbol_mapper_destroy'

    def __init__(self, constellation=None, lib=None):
        _Handle.__init__(self, lib)
        pts, rule = constellation if constellation is not None else qpsk_constellation()
        pts = _c64(pts)
        c = _Constellation(pts.ctypes.data, pts.size, int(rule))
        self._ck(self._dll.gfdm_symbol_mapper_create(byref(self._h), byref(c)))

    def n_points(self):
        return self._dll.gfdm_symbol_mapper_n_points(self._h)

    def bits_per_symbol(self):
        return self._dll.gfdm_symbol_mapper_bits_per_symbol(self._h)

    def decision_rule(self):
        return self._dll.gfdm_symbol_mapper_decision_rule(self._h)

    def points(self):
        out = np.empty(self.n_points(), np.complex64)
        self._ck(self._dll.gfdm_symbol_mapper_points(self._h, _ptr(out)))
        return out

    def map_chunks(self, chunks):
        c = _u8(chunks)
        out = np.empty(c.shape, np.complex64)
        self._ck(self._dll.gfdm_symbol_mapper_map_chunks_batch(self._h, _ptr(out), _ptr(c), c.size, MEM_HOST))
        return out

    def decide(self, symbols):
        a = _c64(symbols)
        out = np.empty(a.shape, np.uint8)
        self._ck(self._dll.gfdm_symbol_mapper_decide_batch(self._h, _ptr(out), _ptr(a), a.size, MEM_HOST))
        return out

    def bits2symbols(self, bits):
        b = _u8(bits).ravel()
        bps = max(self.bits_per_symbol(), 1)
        if b.size % bps:
            raise RuntimeError('number of bits MUST be a multiple of bits_per_symbol(%d)' % bps)
        out = np.empty(b.size // bps, np.complex64)
        self._ck(self._dll.gfdm_symbol_mapper_bits2symbols_batch(self._h, _ptr(out), _ptr(b), out.size, MEM_HOST))
        return out

    def symbols2bits(self, symbols):
        a = _c64(symbols).ravel()
        out = np.empty(a.size * max(self.bits_per_symbol(), 1), np.uint8)
        self._ck(self._dll.gfdm_symbol_mapper_symbols2bits_batch(self._h, _ptr(out), _ptr(a), a.size, MEM_HOST))
        return out

    def map_chunks_ptr(self, out_ptr, chunks_ptr, n):
        self._ck(self._dll.gfdm_symbol_mapper_map_chunks_batch(self._h, c_void_p(out_ptr), c_void_p(chunks_ptr), n,
                                                               MEM_DEVICE))

    def decide_ptr(self, out_ptr, in_ptr, n):
        self._ck(self._dll.gfdm_symbol_mapper_decide_batch(self._h, c_void_p(out_ptr), c_void_p(in_ptr), n,
                                                           MEM_DEVICE))
