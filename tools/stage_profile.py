#!/usr/bin/env python3
"""Per-stage cycle breakdown of the fused kernels (instrumented build: make -C gr-gfdm_b200 prof).
usage (on a GPU box): python tools/stage_profile.py [workload] [frames]"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'gr-gfdm_b200'))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from gfdm_b200 import capi  # noqa: E402
import bench  # noqa: E402

w = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else 'c3'])
frames = int(sys.argv[2]) if len(sys.argv) > 2 else w['frames']
lib = capi.load(os.environ.get('GFDM_PROF_LIB') or os.path.join(ROOT, 'gr-gfdm_b200', 'lib', 'libgfdm_b200_prof.so'))
twopass = w['K'] == 2048
dbg = lib.dll.gfdm_debug_stage_cycles2 if twopass else lib.dll.gfdm_debug_stage_cycles
dbg.argtypes = [ctypes.c_void_p, ctypes.c_int]
tx, rx = bench.make_taps(w)
mod = capi.Modulator(w['M'], w['K'], w['L'], tx, lib=lib)
dem = capi.Demodulator(w['M'], w['K'], w['L'], rx, lib=lib)
d_in = torch.from_numpy(bench.make_symbols(w, frames, 0)).cuda()
d_tx = torch.empty_like(d_in)
d_out = torch.empty_like(d_in)
eq = torch.ones_like(d_in)
names = {0: 'mod: wait bulk load', 1: 'mod: A read staging', 2: 'mod: A fft+row write', 3: 'mod: row fft (warp0)',
         4: 'mod: barrier after rows', 7: 'mod: C read columns', 8: 'mod: barrier behind the reads', 5: 'mod: tail load issue',
         6: 'mod: C table+ifft+store',
         15: 'rx: loop top/store issue', 16: 'rx: wait bulk load', 17: "rx: A' read staging", 18: "rx: A' fft+table+wait store", 19: 'rx: row write',
         20: 'rx: row fft (warp0)', 21: 'rx: barrier after rows', 22: "rx: C' read columns", 23: "rx: C' ifft+staging"}


if twopass:
    names = {0: 'mod2: wait bulk loads', 1: 'mod2: step 0 (read, fft)', 2: 'mod2: step 1 + row writes', 3: 'mod2: row fft (warp0)',
             4: 'mod2: barrier + C read columns', 5: 'mod2: C table+ifft+store', 6: 'mod2p: pass 1 rows out of tensor memory', 7: 'mod2p: barrier after the row FFT', 8: 'mod2p: table loads + column reads',
             9: 'mod2p: C pass 1 (ifft, unpark, stores)',
             16: "rx2: A' four steps + row writes", 17: 'rx2: row fft (warp0)', 18: "rx2: barrier + C' reads",
             19: "rx2: C' ifft+staging+stores (parked: pass 0 ifft+park)", 20: 'rx2: step 0', 21: 'rx2: step 1', 22: 'rx2: step 2', 23: 'rx2: step 3',
             24: 'rx2p: pass 1 rows out of tensor memory', 25: "rx2p: C' pass 1 (ifft, unpark, staging, bulk stores)",
             26: 'rx2p: wait for the last store, issue step 1 of the next frame'}


def run(label, fn, reps=5):
    fn()
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 32)()
    dbg(None, 1)
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dbg(buf, 0)
    v = np.array(list(buf), dtype=np.float64) / (reps * frames)
    tot = v.sum()
    print('%s: %.0f cycles per frame (sum over stages, thread 0 of each CTA)' % (label, tot))
    for i in range(32):
        if v[i] > 0:
            print('   %-28s %8.0f  %5.1f%%' % (names.get(i, str(i)), v[i], 100 * v[i] / tot))


run('modulator', lambda: mod.modulate_ptr(d_tx.data_ptr(), d_in.data_ptr(), frames))
run('receiver', lambda: dem.demodulate_ptr(d_out.data_ptr(), d_tx.data_ptr(), 0, frames))
if not twopass:
    run('receiver+eq', lambda: dem.demodulate_ptr(d_out.data_ptr(), d_tx.data_ptr(), eq.data_ptr(), frames))
