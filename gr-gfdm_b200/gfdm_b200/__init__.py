"""gfdm_b200 -- B200-native GFDM baseband engine (drop-in for gr-gfdm's kernel layer).

Host-side Python mirror of the reference's ``gfdm_python`` bindings on top of the
C ABI in include/gfdm_b200.h.  The compute path is hand-written sm_100a CUDA in
``gr-gfdm_b200/csrc`` built into ``gr-gfdm_b200/lib/libgfdm_b200.so``; there is no
CPU fallback -- importing the kernel classes without that library raises.
"""
from . import capi  # noqa: F401
from .capi import (Advanced_receiver, Cyclic_prefixer, Demodulator, FFT, Modulator,  # noqa: F401
                   Preamble_channel_estimator, Resource_mapper, Transmitter, load,
                   qpsk_constellation)

__all__ = ['capi', 'load', 'FFT', 'Modulator', 'Demodulator', 'Advanced_receiver',
           'Resource_mapper', 'Cyclic_prefixer', 'Preamble_channel_estimator', 'Transmitter',
           'qpsk_constellation']
