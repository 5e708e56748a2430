#!/usr/bin/env python3
"""End-to-end (HOST buffers through the C ABI) throughput of the C3 chain against the pipeline chunk size.
usage (GPU box): for mb in 4 8 16 32 64 128; do GFDM_PIPE_CHUNK_MB=$mb python tools/e2e_sweep.py; done"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'gr-gfdm_b200'))
import torch  # noqa: E402
from gfdm_b200 import capi, design  # noqa: E402

M, K, L, frames = 15, 1024, 2, 4096
lib = capi.load()
lib.set_device(0)
tx = design.get_frequency_domain_filter('rrc', .5, M, K, L).astype(np.complex64)
rx = design.get_zero_forcing_taps('rrc', .5, M, K, L).astype(np.complex64)
mod, dem = capi.Modulator(M, K, L, tx, lib=lib), capi.Demodulator(M, K, L, rx, lib=lib)
rng = np.random.default_rng(1)
host_in = torch.from_numpy(design.get_random_qam16(frames * M * K, rng).reshape(frames, -1).astype(np.complex64)).pin_memory()
host_tx = torch.empty_like(host_in).pin_memory()
host_out = torch.empty_like(host_in).pin_memory()


def step():
    mod.modulate_batch_host_ptr(host_tx.data_ptr(), host_in.data_ptr(), frames)
    dem.demodulate_batch_host_ptr(host_out.data_ptr(), host_tx.data_ptr(), 0, frames)


step()
torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    t0 = time.perf_counter()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    best = min(best, (time.perf_counter() - t0) / 3)
print(json.dumps({'chunk_mb': os.environ.get('GFDM_PIPE_CHUNK_MB', 'default(32)'), 'ms_per_step': best * 1e3,
                  'frames_per_s': frames / best, 'gbs_each_way': 2 * frames * M * K * 8 / best / 1e9}))
