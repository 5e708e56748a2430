#!/usr/bin/env python3
"""bench.py -- GFDM hot-path throughput on B200 (BASELINE.json metric) next to the reference's CPU kernels.

  python bench.py --gpus N --steps K --warmup W                   # this repo's CUDA path, headline workload (c3)
  python bench.py --impl reference --gpus N --steps K ...         # the reference's CPU kernels (oracle/_ref)
  python bench.py --workload c2|c4|c5|c1 ...                      # the other BASELINE.json configs, as named there
  python bench.py --workload c5 --sweep                           # configs[4]: 2^10 .. 2^20 frames, sharded over the ranks

Workloads (BASELINE.json `configs`, sizes of SURVEY.md section 8; one *step* = one pass of the chain over the batch):
  c3 (default, the configuration the metric is quoted on)  K=1024 M=15 L=2, 16-QAM: modulator_kernel_cc::generic_work
      on every frame, then receiver_kernel_cc::generic_work on every modulated frame; 4096 frames per GPU.
  c1  K=16 M=5 (the reference's qa shape), same chain.
  c2  K=64 M=9 A=52 cp=16 cs=8: transmitter_kernel::generic_work (mapper + modulator + preamble + CP + window, ONE
      kernel) -> receiver_kernel_cc::generic_work_equalize (per-frame channel) reading the block out of every frame in
      place (remove_prefix_cc fused into the receiver's loads).
  c4  K=256 M=15 A=208: preamble_channel_estimator_cc::estimate_frame on the received preamble of every frame, then
      advanced_receiver_kernel_cc::generic_work_equalize with that estimate, 4 SIC iterations, per-frame multipath + AWGN.
  c5  K=2048 M=15: modulate -> demodulate with the two-pass kernels; --sweep runs 2^10 .. 2^20 total frames.

One process per GPU (torchrun for N > 1); frames are independent, so ranks shard the batch and never communicate on the
data path ("scaling": "weak").  Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max
over ranks.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, 'gr-gfdm_b200'))
os.environ.setdefault('HOME', '/tmp')

import numpy as np  # noqa: E402

WORKLOADS = {
    'c3': dict(kind='modem', K=1024, M=15, L=2, frames=4096, alpha=0.5, seed=1003,
               desc='K=1024 M=15 L=2 16-QAM modulate->ZF-demodulate, 4096 frames/GPU (BASELINE configs[2])'),
    'c1': dict(kind='modem', K=16, M=5, L=2, frames=1 << 18, alpha=0.5, seed=1001,
               desc='K=16 M=5 L=2 modulate->demodulate (BASELINE configs[0])'),
    'c2': dict(kind='txrx', K=64, M=9, L=2, A=52, cp=16, cs=8, frames=1 << 16, alpha=0.2, seed=1002,
               desc='K=64 M=9 L=2 A=52 cp=16 cs=8 transmitter_kernel chain -> equalising receiver on the framed samples '
                    '(BASELINE configs[1])'),
    'c4': dict(kind='est_sic', K=256, M=15, L=2, A=208, frames=1 << 14, alpha=0.5, seed=1004, ic_iter=4, snr_db=20.0,
               desc='K=256 M=15 L=2 A=208 QPSK preamble_channel_estimator -> advanced receiver (4 SIC iterations), '
                    'per-frame 4..8-tap multipath + AWGN 20 dB (BASELINE configs[3])'),
    'c5': dict(kind='modem', K=2048, M=15, L=2, frames=2048, alpha=0.5, seed=1005,
               desc='K=2048 M=15 L=2 modulate->demodulate, two-pass kernels (BASELINE configs[4] shape)'),
}
RX_TAPS_LABEL = 'canonical dual (ZF) window of the transmit pulse truncated to overlap 2 (design.get_zero_forcing_taps_full, ' \
                'rx_overlap=2): the full-width ZF needs overlap = K and costs the GPU kernels nothing, the CPU reference K/2 x'


def dist_env():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


def make_taps(w):
    from gfdm_b200 import design
    tx = design.get_frequency_domain_filter('rrc', w['alpha'], w['M'], w['K'], w['L'])
    rx = design.get_zero_forcing_taps_full(tx, w['M'], w['K'], w['L'], rx_overlap=w['L'])
    return tx.astype(np.complex64), rx.astype(np.complex64)


def qam16(rng, shape):
    from gfdm_b200 import design
    pts = design.qam16_points().astype(np.complex64)
    out = np.empty(shape, np.complex64)
    flat = out.reshape(shape[0], -1)
    step = max(1, (1 << 24) // flat.shape[1])
    for f0 in range(0, shape[0], step):
        f1 = min(shape[0], f0 + step)
        flat[f0:f1] = pts[rng.integers(0, 16, (f1 - f0, flat.shape[1]))]
    return out


def make_symbols(w, n_frames, rank):
    return qam16(np.random.default_rng(w['seed'] + rank), (n_frames, w['K'] * w['M']))


# ------------------------------------------------------------------------------------------------------------
# The chains.  A chain owns its handles on ONE library (the CUDA product or a CPU oracle) and the host-side inputs of
# its shard; `host_pass(bufs)` runs the chain through HOST pointers (the CPU baseline and the end-to-end leg).
class ModemChain(object):
    kind = 'modem'

    def __init__(self, w, lib):
        from gfdm_b200 import capi
        self.w, self.lib = w, lib
        self.M, self.K, self.L = w['M'], w['K'], w['L']
        self.N = self.M * self.K
        tx, rx = make_taps(w)
        self.mod = capi.Modulator(self.M, self.K, self.L, tx, lib=lib)
        self.dem = capi.Demodulator(self.M, self.K, self.L, rx, lib=lib)
        self.handles = [self.mod, self.dem]
        # (stage name, algorithmic bytes per frame): SURVEY 8d / BASELINE.md section 3
        self.stages = [('modulator', 16 * self.N), ('receiver', 16 * self.N)]
        self.baseline_bytes = 32 * self.N

    def host_inputs(self, n, rank):
        return {'sym': make_symbols(self.w, n, rank)}

    def host_buffers(self, n):
        return {'tx': (n, self.N), 'out': (n, self.N)}

    def host_pass(self, p, n):  # p: dict of raw host pointers
        self.mod.modulate_batch_host_ptr(p['tx'], p['sym'], n)
        self.dem.demodulate_batch_host_ptr(p['out'], p['tx'], 0, n)

    def host_bytes(self, n):  # (h2d, d2h) per pass
        return 2 * n * self.N * 8, 2 * n * self.N * 8

    host_path = 'gfdm_modulator_work_batch + gfdm_receiver_work_batch, GFDM_MEM_HOST, pinned host buffers'

    def device_setup(self, torch, dev, inputs, n):
        self.d_in = torch.from_numpy(inputs['sym']).to(dev)
        self.d_tx = torch.empty_like(self.d_in)
        self.d_out = torch.empty_like(self.d_in)
        self.n = n

    def device_stage(self, i, n=None):
        n = self.n if n is None else n
        if i == 0:
            self.mod.modulate_ptr(self.d_tx.data_ptr(), self.d_in.data_ptr(), n)
        else:
            self.dem.demodulate_ptr(self.d_out.data_ptr(), self.d_tx.data_ptr(), 0, n)

    def kernel_names(self):
        return {'modulator': self.mod.last_kernel(), 'receiver': self.dem.last_kernel()}

    def check(self):
        chk = self.d_out[:2].cpu().numpy()
        assert np.isfinite(chk).all(), 'non-finite demodulator output'


class TxRxChain(object):
    """BASELINE configs[1]: transmitter_kernel full chain, then the receive side up to the equalising receiver."""
    kind = 'txrx'

    def __init__(self, w, lib):
        from gfdm_b200 import capi, design
        self.w, self.lib = w, lib
        M, K, L, A, cp, cs = w['M'], w['K'], w['L'], w['A'], w['cp'], w['cs']
        self.M, self.K, self.L, self.A, self.N = M, K, L, A, M * K
        ramp = cs
        cfg = design.get_gfdm_configuration(M, K, A, L, cp, cs, 'rrc', w['alpha'])
        tx, rx = make_taps(w)
        self.P = cfg.preamble_len
        self.W = self.N + cp + cs
        self.os = self.P + self.W
        self.n_in = A * M
        self.tx = capi.Transmitter(M, K, A, cp, cs, ramp, cfg.subcarrier_map, True, L, tx, cfg.window_taps, cfg.cyclic_shifts,
                                   cfg.full_preambles, lib=lib)
        self.rx = capi.Demodulator(M, K, L, rx, lib=lib)
        self.off = self.P + cp                                           # first sample of the block inside a frame
        self.handles = [self.tx, self.rx]
        self.stages = [('transmitter', 8 * (self.n_in + self.os)), ('receiver_equalize', 24 * self.N)]
        self.baseline_bytes = 8 * (self.n_in + self.os) + 24 * self.N   # 9,760 + 13,824 at C2

    def host_inputs(self, n, rank):
        rng = np.random.default_rng(self.w['seed'] + rank)
        return {'sym': qam16(rng, (n, self.n_in)), 'eq': np.ones((n, self.N), np.complex64)}

    def host_buffers(self, n):
        return {'frame': (n, self.os), 'out': (n, self.N)}

    def host_pass(self, p, n):
        from gfdm_b200 import capi
        self.tx.work_ptr(p['frame'], p['sym'], self.n_in, n, mem=capi.MEM_HOST)
        self.rx.demodulate_strided_ptr(p['out'], p['frame'], p['eq'], self.os, self.off, n, mem=capi.MEM_HOST)

    def host_bytes(self, n):
        return 8 * n * (self.n_in + self.os + self.N), 8 * n * (self.os + self.N)

    host_path = 'gfdm_transmitter_work_batch + gfdm_receiver_work_strided_batch(eq), GFDM_MEM_HOST'

    def device_setup(self, torch, dev, inputs, n):
        self.d_sym = torch.from_numpy(inputs['sym']).to(dev)
        self.d_eq = torch.from_numpy(inputs['eq']).to(dev)
        self.d_frame = torch.empty((n, self.os), dtype=torch.complex64, device=dev)
        self.d_out = torch.empty((n, self.N), dtype=torch.complex64, device=dev)
        self.n = n

    def device_stage(self, i, n=None):
        n = self.n if n is None else n
        if i == 0:
            self.tx.work_ptr(self.d_frame.data_ptr(), self.d_sym.data_ptr(), self.n_in, n)
        else:
            self.rx.demodulate_strided_ptr(self.d_out.data_ptr(), self.d_frame.data_ptr(), self.d_eq.data_ptr(), self.os, self.off, n)

    def kernel_names(self):
        return {'transmitter': self.tx.last_kernel(), 'receiver_equalize': self.rx.last_kernel()}

    def check(self):
        assert np.isfinite(self.d_out[:2].cpu().numpy()).all(), 'non-finite receiver output'


class EstSicChain(object):
    """BASELINE configs[3]: channel estimate from the received preamble, then the interference-cancelling receiver
    equalising with that estimate; per-frame multipath channel + AWGN."""
    kind = 'est_sic'

    def __init__(self, w, lib):
        from gfdm_b200 import capi, design
        self.w, self.lib = w, lib
        M, K, L, A = w['M'], w['K'], w['L'], w['A']
        self.M, self.K, self.L, self.A, self.N = M, K, L, A, M * K
        self.smap = design.get_subcarrier_map(K, A, dc_free=True)
        tx = design.get_frequency_domain_filter('rrc', w['alpha'], M, K, L).astype(np.complex64)
        self.tx_taps = tx
        _, self.core = design.mapped_preamble(design.PREAMBLE_SEED, 'rrc', w['alpha'], A, K, self.smap, L, K // 4, K // 8,
                                              use_zadoff_chu=True)
        self.core = self.core.astype(np.complex64)
        self.est = capi.Preamble_channel_estimator(M, K, A, True, 1, self.core, lib=lib)
        self.adv = capi.Advanced_receiver(M, K, L, np.conj(tx), self.smap, w['ic_iter'], capi.qpsk_constellation(), 0, lib=lib)
        self.handles = [self.est, self.adv]
        self.stages = [('estimator', 8 * (2 * K + self.N)), ('advanced_receiver', 24 * self.N)]
        self.baseline_bytes = 8 * (2 * K + self.N) + 24 * self.N        # 34,816 + 92,160 at C4

    def host_inputs(self, n, rank):
        """Received preambles [n][2K] and frames [n][N]: QPSK on the active subcarriers -> mapper -> modulator (this chain's
        own library), per-frame 4..8-tap complex Gaussian channel with exponential power profile and unit energy applied by
        circular convolution, AWGN at snr_db."""
        from gfdm_b200 import capi, design
        M, K, N, A = self.M, self.K, self.N, self.A
        rng = np.random.default_rng(self.w['seed'] + rank)
        mp = capi.Resource_mapper(M, K, A, self.smap, True, lib=self.lib)
        mod = capi.Modulator(M, K, self.L, self.tx_taps, lib=self.lib)
        pts, _ = capi.qpsk_constellation()
        pre = np.empty((n, 2 * K), np.complex64)
        frm = np.empty((n, N), np.complex64)
        sigma = 10.0 ** (-self.w['snr_db'] / 20.0)
        blk = max(1, (1 << 22) // N)
        for f0 in range(0, n, blk):
            f1 = min(n, f0 + blk)
            nb = f1 - f0
            sym = pts[rng.integers(0, 4, (nb, A * M))]
            x = mod.modulate_batch(mp.map_to_resources_batch(sym))
            ntap = rng.integers(4, 9, nb)
            h = (rng.standard_normal((nb, 8)) + 1j * rng.standard_normal((nb, 8))) * np.exp(-0.5 * np.arange(8))[None, :]
            h[np.arange(8)[None, :] >= ntap[:, None]] = 0
            h /= np.sqrt(np.sum(np.abs(h) ** 2, axis=1, keepdims=True))
            y = np.fft.ifft(np.fft.fft(x, axis=1) * np.fft.fft(h, N, axis=1), axis=1)
            p = np.fft.ifft(np.fft.fft(self.core.reshape(1, 2, K), axis=2) * np.fft.fft(h, K, axis=1)[:, None, :], axis=2).reshape(nb, 2 * K)
            sx, sp = np.sqrt(np.mean(np.abs(y) ** 2)), np.sqrt(np.mean(np.abs(p) ** 2))
            frm[f0:f1] = y + sigma * sx / np.sqrt(2) * (rng.standard_normal(y.shape) + 1j * rng.standard_normal(y.shape))
            pre[f0:f1] = p + sigma * sp / np.sqrt(2) * (rng.standard_normal(p.shape) + 1j * rng.standard_normal(p.shape))
        return {'pre': pre, 'frm': frm}

    def host_buffers(self, n):
        return {'h': (n, self.N), 'out': (n, self.N)}

    def host_pass(self, p, n):
        from gfdm_b200 import capi
        self.est.estimate_frame_ptr(p['h'], p['pre'], n, mem=capi.MEM_HOST)
        self.adv.demodulate_ptr(p['out'], p['frm'], p['h'], n, mem=capi.MEM_HOST)

    def host_bytes(self, n):
        return 8 * n * (2 * self.K + 2 * self.N), 8 * n * (2 * self.N)

    host_path = 'gfdm_channel_estimator_estimate_frame_batch + gfdm_advanced_receiver_work_batch(eq), GFDM_MEM_HOST'

    def device_setup(self, torch, dev, inputs, n):
        self.d_pre = torch.from_numpy(inputs['pre']).to(dev)
        self.d_frm = torch.from_numpy(inputs['frm']).to(dev)
        self.d_h = torch.empty_like(self.d_frm)
        self.d_out = torch.empty_like(self.d_frm)
        self.n = n

    def device_stage(self, i, n=None):
        n = self.n if n is None else n
        if i == 0:
            self.est.estimate_frame_ptr(self.d_h.data_ptr(), self.d_pre.data_ptr(), n)
        else:
            self.adv.demodulate_ptr(self.d_out.data_ptr(), self.d_frm.data_ptr(), self.d_h.data_ptr(), n)

    def kernel_names(self):
        return {'estimator': self.est.last_kernel(), 'advanced_receiver': self.adv.last_kernel()}

    def check(self):
        assert np.isfinite(self.d_out[:2].cpu().numpy()).all(), 'non-finite receiver output'


CHAINS = {'modem': ModemChain, 'txrx': TxRxChain, 'est_sic': EstSicChain}


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def count(self):
        return len(self.lines) if self.proc is not None else 1 << 30

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for l in self.lines:
            p = [x.strip() for x in l.split(',')]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smax.append(float(p[2]))
                power.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(names, p[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': sorted(reasons)}


def cpu_reference_run(w, seconds_budget, max_threads=None):
    """Time the reference's own CPU kernels (oracle/_ref = unmodified gr-gfdm sources + scalar FFTW/VOLK shims; falls
    back to the plain-C port) on the host cores: one chain instance per core, each running the SAME chain as the GPU arm
    over its own `frames_per_thread` frames of the same synthetic workload."""
    from gfdm_b200 import capi
    ref_so = os.path.join(ROOT, 'oracle', '_ref', 'libgfdm_ref.so')
    port_so = os.path.join(ROOT, 'oracle', '_ref', 'libgfdm_port.so')
    if os.path.exists(ref_so):
        lib, kind = capi.load(ref_so), 'reference'
    else:
        if not os.path.exists(port_so):
            subprocess.run(['make', '-C', os.path.join(ROOT, 'oracle'), 'port'], check=True, stdout=subprocess.DEVNULL)
        lib, kind = capi.load(port_so), 'port'
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    if max_threads:
        cores = min(cores, max_threads)
    Chain = CHAINS[w['kind']]
    # calibrate on one thread, then size the sample so that all threads run ~seconds_budget
    probe = Chain(w, lib)

    def buffers(chain, n, inputs):
        arrs = dict(inputs)
        for name, shape in chain.host_buffers(n).items():
            arrs[name] = np.empty(shape, np.complex64)
        return arrs, {k: v.ctypes.data for k, v in arrs.items()}

    pin = probe.host_inputs(2, 0)
    parr, pptr = buffers(probe, 2, pin)
    t0 = time.perf_counter()
    probe.host_pass(pptr, 2)
    per_frame = (time.perf_counter() - t0) / 2
    fpt = int(max(2, min(4096, seconds_budget / max(per_frame, 1e-7))))
    chains = [probe] + [Chain(w, lib) for _ in range(cores - 1)]
    inputs = probe.host_inputs(fpt, 0)
    bufs = [buffers(c, fpt, inputs) for c in chains]  # inputs shared (read-only), outputs per thread

    def work(i):
        chains[i].host_pass(bufs[i][1], fpt)   # ctypes releases the GIL

    def run_all():
        th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0

    return dict(kind=kind, cores=cores, frames_per_thread=fpt, run=run_all, N=w['K'] * w['M'],
                sample='%d threads x %d frames of the same workload (%s), %s kernels%s' % (
                    cores, fpt, probe.host_path.split(',')[0], 'reference gr-gfdm' if kind == 'reference' else 'oracle port',
                    ' + scalar FFTW/VOLK shims' if kind == 'reference' else ''))


METRIC = {'modem': 'GFDM frames/s (mod+demod)', 'txrx': 'GFDM frames/s (transmitter chain + equalising receiver)',
          'est_sic': 'GFDM frames/s (channel estimate + SIC receiver)'}


def run_reference(args, w):
    rank, _, world = dist_env()
    if rank != 0:
        return
    ref = cpu_reference_run(w, seconds_budget=max(1.0, args.cpu_seconds / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        ref['run']()
    times = [ref['run']() for _ in range(args.steps)]
    t = float(np.mean(times))
    frames = ref['cores'] * ref['frames_per_thread']
    value = frames / t
    line = {
        'impl': 'reference', 'metric': METRIC[w['kind']], 'value': value, 'unit': 'frames/s',
        'msamples_per_s': value * ref['N'] / 1e6, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': t * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'complex64 (fp32)',
        'data': 'synthetic', 'config': {'workload': w['desc'], 'frames_per_step': frames, 'device': 'host CPU'},
        'cpu_baseline': {'value': value, 'unit': 'frames/s', 'cores': ref['cores'], 'kind': ref['kind'], 'sample': ref['sample']},
        'e2e': {'value': value, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(torch, local_rank):
    """Pin this rank to the CPUs next to its GPU (sysfs local_cpulist) BEFORE the pinned host buffers are
    allocated, so that first-touch places them on the GPU's NUMA node and the HOST-buffer legs of N ranks do not
    share one memory controller.  Returns (previous affinity, description); a no-op when sysfs has no answer."""
    try:
        prev = os.sched_getaffinity(0)
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = '%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        txt = open('/sys/bus/pci/devices/%s/local_cpulist' % bdf).read().strip()
        cpus = set()
        for part in txt.split(','):
            if not part:
                continue
            lo, _, hi = part.partition('-')
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= prev
        if cpus and cpus != prev:
            os.sched_setaffinity(0, cpus)
            return prev, '%s: %d of %d cpus' % (bdf, len(cpus), len(prev))
        return prev, '%s: all %d cpus local' % (bdf, len(prev))
    except Exception as e:  # noqa: BLE001 -- containers without sysfs, odd topologies
        return None, 'unavailable (%s)' % type(e).__name__


def measure_link_ceiling(torch, dist, dev, world, barrier, mb=256, chunk_mb=32, secs=0.4):
    """What the box can copy between pinned host memory and its GPUs while ALL ranks copy at once: plain
    cudaMemcpyAsync (torch copy_) in 32 MB chunks, host->device and device->host alone and together, no kernels.  This
    is the ceiling of every HOST-buffer entry of the library; tools/pcie_ceiling.cu measures the same without torch."""
    n = mb << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.fill_(1)
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.zeros(n, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    c = chunk_mb << 20
    res = {}
    for name, dirs in (('h2d', (True, False)), ('d2h', (False, True)), ('h2d+d2h', (True, True))):
        def one_pass():
            for o in range(0, n, c):
                if dirs[0]:
                    with torch.cuda.stream(s_in):
                        d_in[o:o + c].copy_(h_in[o:o + c], non_blocking=True)
                if dirs[1]:
                    with torch.cuda.stream(s_out):
                        h_out[o:o + c].copy_(d_out[o:o + c], non_blocking=True)
        one_pass()
        barrier()
        t0 = time.perf_counter()
        passes = 0
        while True:
            one_pass()
            s_in.synchronize()
            s_out.synchronize()
            passes += 1
            dt = time.perf_counter() - t0
            if dt >= secs:
                break
        gbs = passes * n * (int(dirs[0]) + int(dirs[1])) / dt / 1e9
        t = torch.tensor([gbs, -gbs], dtype=torch.float64, device=dev)
        if world > 1:
            tsum = t.clone()
            dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            res[name] = {'aggregate_gbs': float(tsum[0].item()), 'slowest_rank_gbs': float(-tmax[1].item())}
        else:
            res[name] = {'aggregate_gbs': gbs, 'slowest_rank_gbs': gbs}
        barrier()
    res['how'] = 'pinned cudaMemcpyAsync, %d MB per direction per pass in %d MB chunks, every rank at once, %.1f s' % (mb, chunk_mb, secs)
    return res


def timed_steps(torch, stream, chain, n_steps, barrier, n_frames=None, loops=1):
    """K timed steps; per-stage CUDA events.  `loops`: the step runs the chain `loops` times over the resident buffers
    (batches larger than what is kept resident).  Returns (total ms, [per-stage mean ms])."""
    ns = len(chain.stages)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(ns + 1)] for _ in range(n_steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        t_start.record(stream)
        for i in range(n_steps):
            for lp in range(loops):
                first = lp == 0
                for s in range(ns):
                    if first:
                        ev[i][s].record(stream)
                    chain.device_stage(s, n_frames)
                if first:
                    ev[i][ns].record(stream)
        t_end.record(stream)
    barrier()
    total = t_start.elapsed_time(t_end)
    stage_ms = [float(np.mean([e[s].elapsed_time(e[s + 1]) for e in ev])) for s in range(ns)]
    return total, stage_ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c3', choices=sorted(WORKLOADS))
    ap.add_argument('--frames', type=int, default=0, help='frames per GPU (default: workload value)')
    ap.add_argument('--sweep', action='store_true', help='frame-batch sweep 2^10 .. 2^20 TOTAL frames (BASELINE configs[4])')
    ap.add_argument('--sweep-max-log2', type=int, default=20)
    ap.add_argument('--cpu-seconds', type=float, default=12.0, help='CPU baseline budget (seconds of wall clock)')
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-latency', action='store_true')
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.frames:
        w['frames'] = args.frames
    if args.impl == 'reference':
        run_reference(args, w)
        return

    import torch
    import torch.distributed as dist
    from gfdm_b200 import capi, design
    from gfdm_b200.sharding import shard_bounds

    rank, local_rank, world = dist_env()
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = capi.load()
    lib.set_device(local_rank)

    K, M, L = w['K'], w['M'], w['L']
    N = K * M
    sweep_totals = [1 << e for e in range(10, args.sweep_max_log2 + 1)] if args.sweep else []
    resident_cap = max(1, (16 << 30) // (8 * N))   # frames kept resident per buffer (16 GB); larger batches loop over them
    if args.sweep:
        lo, hi = shard_bounds(sweep_totals[-1], world, rank)
        w['frames'] = min(hi - lo, resident_cap)
    frames = w['frames']
    chain = CHAINS[w['kind']](w, lib)
    stream = torch.cuda.Stream(device=dev)
    for h in chain.handles:
        h.set_stream(stream.cuda_stream)

    prev_affinity, affinity_desc = bind_to_gpu_numa_node(torch, local_rank)
    inputs = chain.host_inputs(frames, rank)
    chain.device_setup(torch, dev, inputs, frames)
    n_stages = len(chain.stages)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(n=None):
        for s in range(n_stages):
            chain.device_stage(s, n)

    def launch_total():
        return sum(h.launch_count() for h in chain.handles)

    # nvidia-smi reports every 100 ms and takes a while to start, the timed region lasts milliseconds: the sampler
    # runs from before the warm-up, the same kernel loop keeps the GPU loaded until a first sample exists, the timed
    # steps follow immediately and the loop continues untimed until a few more samples are in -- so the samples
    # bracket the timed region under identical load
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_w = time.perf_counter()
    with torch.cuda.stream(stream):
        n_w = 0
        while n_w < max(args.warmup, 3) or (sampler.count() < 1 and time.perf_counter() - t_w < 3.0):
            step()
            n_w += 1
            if n_w % 8 == 0:
                stream.synchronize()
    barrier()
    launches0 = launch_total()
    total_ms, stage_ms = timed_steps(torch, stream, chain, args.steps, barrier)
    launches1 = launch_total()
    kernel_names = chain.kernel_names()

    # ---- frame-batch sweep (BASELINE configs[4]): total frames 2^10 .. 2^20 sharded over the ranks -----------------
    sweep = None
    if args.sweep:
        sweep = []
        for total in sweep_totals:
            lo, hi = shard_bounds(total, world, rank)
            mine = hi - lo
            nres = min(mine, frames)
            loops = (mine + nres - 1) // nres if mine else 0
            ksteps = max(3, min(args.steps, (1 << 16) // max(1, mine)))
            if mine:
                with torch.cuda.stream(stream):
                    step(nres)
            tms, sms = timed_steps(torch, stream, chain, ksteps, barrier, nres, max(loops, 1)) if mine else (0.0, [0.0] * n_stages)
            t = torch.tensor([tms / ksteps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            done = nres * loops   # frames this rank processed per step (a multiple of the resident batch)
            tot = torch.tensor([float(done)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            processed = float(tot.item())
            sweep.append({'total_frames': total, 'frames_processed_per_step': int(processed), 'frames_per_gpu': mine,
                          'resident_frames_per_launch': nres, 'launch_loops': loops, 'steps': ksteps, 'ms_per_step': ms,
                          'frames_per_s': processed / (ms * 1e-3),
                          'chain_frac_of_hbm_peak': None})
    t_w = time.perf_counter()
    n0 = sampler.count()
    with torch.cuda.stream(stream):
        while sampler.count() < n0 + 2 and time.perf_counter() - t_w < 0.6:
            for _ in range(8):
                step()
            stream.synchronize()
    clocks = sampler.stop()
    clocks['how'] = 'same kernel loop running before, during and after the timed region (nvidia-smi period 100 ms)'
    launches = launches1 - launches0
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    value = world * frames / (ms_per_step * 1e-3)
    chain.check()

    # ---- end to end through the C ABI with HOST buffers (H2D + D2H inside the timed region) -------------------------
    e2e = None
    e2e_variants = {}
    link = None
    if not args.no_e2e and not args.sweep:
        def pinned(a):
            return torch.from_numpy(a).pin_memory()

        def timed_host(fn, steps):
            fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                fn()
            barrier()
            dt = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return float(dt.item())

        link = measure_link_ceiling(torch, dist, dev, world, barrier)
        h_in = {k: pinned(v) for k, v in inputs.items()}
        h_buf = {k: torch.empty(shape, dtype=torch.complex64).pin_memory() for k, shape in chain.host_buffers(frames).items()}
        ptrs = {k: v.data_ptr() for k, v in list(h_in.items()) + list(h_buf.items())}
        dt = timed_host(lambda: chain.host_pass(ptrs, frames), args.e2e_steps)
        hb, db = chain.host_bytes(frames)

        def variant(dt, h2d, d2h, path):
            rate = world * (h2d + d2h) / dt / 1e9
            return {'value': world * frames / dt, 'unit': 'frames/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'steps': args.e2e_steps, 'ms_per_step': dt * 1e3, 'path': path, 'link_gbs': rate,
                    'frac_of_link_ceiling': rate / link['h2d+d2h']['aggregate_gbs']}

        e2e_variants['cf32'] = variant(dt, hb, db, chain.host_path)
        e2e = e2e_variants['cf32']
        if chain.kind == 'modem' and N % 16 == 0:
            # The same chain with the symbol side as one byte per symbol (chunks in, hard decisions out: SURVEY 8f rank 2) and
            # the two handles software-pipelined over slices of the batch with GFDM_MEM_HOST_ASYNC, so that the receiver's
            # host->device leg overlaps the modulator's device->host leg (both PCIe directions busy).
            sm = capi.Symbol_mapper((design.qam16_points(), capi.DECISION_NEAREST), lib=lib)
            g = torch.Generator().manual_seed(w['seed'] + rank)
            h_ch = torch.randint(0, 16, (frames, N), dtype=torch.uint8, generator=g).pin_memory()
            h_dec = torch.empty_like(h_ch).pin_memory()
            h_tx = h_buf['tx']
            h_iq = torch.empty((frames, N, 2), dtype=torch.int16).pin_memory()
            S = 8 if frames >= 64 else 1
            bounds = [shard_bounds(frames, S, i) for i in range(S)]
            mod, dem = chain.mod, chain.dem
            A_ = capi.MEM_HOST_ASYNC

            def pipelined(mod_call, dem_call):
                def run():
                    mod_call(*bounds[0])
                    for i in range(S):
                        mod.sync()                      # slice i of the samples is in host memory
                        if i + 1 < S:
                            mod_call(*bounds[i + 1])    # ... its successor is modulated while slice i is received
                        dem_call(*bounds[i])
                    dem.sync()
                return run

            def mod_chunks(a, b):
                mod.modulate_chunks_batch_host_ptr(sm, h_tx[a:b].data_ptr(), h_ch[a:b].data_ptr(), b - a, mem=A_)

            def dem_decide(a, b):
                dem.demodulate_decide_batch_host_ptr(sm, h_dec[a:b].data_ptr(), h_tx[a:b].data_ptr(), 0, b - a, mem=A_)

            dt = timed_host(pipelined(mod_chunks, dem_decide), args.e2e_steps)
            e2e_variants['chunks'] = variant(dt, frames * 9 * N, frames * 9 * N,
                                             'gfdm_modulator_work_chunks_batch + gfdm_receiver_work_decide_batch, '
                                             'GFDM_MEM_HOST_ASYNC, %d slices pipelined across the two handles' % S)
            ref_dec = h_dec.clone()
            scale = 8192.0

            def mod_sc16(a, b):
                mod.modulate_chunks_sc16_ptr(sm, h_iq[a:b].data_ptr(), h_ch[a:b].data_ptr(), scale, b - a, mem=A_)

            def dem_sc16(a, b):
                dem.demodulate_decide_sc16_ptr(sm, h_dec[a:b].data_ptr(), h_iq[a:b].data_ptr(), 0, scale, b - a, mem=A_)

            dt = timed_host(pipelined(mod_sc16, dem_sc16), args.e2e_steps)
            e2e_variants['chunks_sc16'] = variant(dt, frames * 5 * N, frames * 5 * N,
                                                  'gfdm_modulator_work_chunks_batch_sc16 + gfdm_receiver_work_decide_batch_sc16 '
                                                  '(int16 I/Q samples on the host side, scale 8192), GFDM_MEM_HOST_ASYNC, %d slices' % S)
            e2e_variants['chunks_sc16']['decisions_differing_from_complex64_samples'] = float((h_dec != ref_dec).float().mean().item())

            def mod_cf(a, b):
                mod.modulate_batch_host_ptr(h_tx[a:b].data_ptr(), h_in['sym'][a:b].data_ptr(), b - a, mem=A_)

            def dem_cf(a, b):
                dem.demodulate_batch_host_ptr(h_buf['out'][a:b].data_ptr(), h_tx[a:b].data_ptr(), 0, b - a, mem=A_)

            dt = timed_host(pipelined(mod_cf, dem_cf), args.e2e_steps)
            e2e_variants['cf32_pipelined'] = variant(dt, hb, db, chain.host_path.replace('GFDM_MEM_HOST', 'GFDM_MEM_HOST_ASYNC') +
                                                     ', %d slices pipelined across the two handles' % S)
            # headline end-to-end number: the byte-wide entries (the caller of a 16-QAM modem has bits), complex64 samples
            e2e = e2e_variants['chunks']
    # ---- latency of ONE frame through the per-frame entry points (host pointers, synchronous: what a block's work()
    # calls frame by frame), pageable NumPy buffers as a GNU Radio buffer would be
    latency = None
    if not args.no_latency and not args.sweep and rank == 0 and chain.kind == 'modem':
        one = np.ascontiguousarray(inputs['sym'][0])
        out1, out2 = np.empty_like(one), np.empty_like(one)
        lat = {}
        for name, fn in (('modulator', lambda: chain.mod.modulate_batch_host_ptr(out1.ctypes.data, one.ctypes.data, 1)),
                         ('receiver', lambda: chain.dem.demodulate_batch_host_ptr(out2.ctypes.data, out1.ctypes.data, 0, 1))):
            for _ in range(50):
                fn()
            ts = []
            for _ in range(1000):
                t0 = time.perf_counter()
                fn()
                ts.append(time.perf_counter() - t0)
            ts = np.array(ts) * 1e6
            lat[name] = {'p50_us': float(np.percentile(ts, 50)), 'p99_us': float(np.percentile(ts, 99)), 'calls': 1000}
        latency = lat
        latency['what'] = 'gfdm_modulator_work / gfdm_receiver_work: one frame, host pointers, synchronous'
    if prev_affinity is not None:
        os.sched_setaffinity(0, prev_affinity)  # the CPU baseline leg uses every host core

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'MEASURED_PEAKS.json hbm_gbs (measured copy)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s (B200_PROFILING.md)'
    # dominant kernel = the stage with the longest launch; its own algorithmic bytes over its own duration
    dom = int(np.argmax(stage_ms))
    dom_name, dom_bytes = chain.stages[dom]
    alg_bytes = float(dom_bytes) * frames
    achieved = alg_bytes / (stage_ms[dom] * 1e-3) / 1e9
    chain_ms = float(sum(stage_ms))
    chain_gbs = chain.baseline_bytes * frames / (chain_ms * 1e-3) / 1e9
    kname = kernel_names[dom_name]
    traffic, traffic_src = None, None
    try:
        tdb = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        ent = tdb.get(kname)
        if ent and int(ent['frames']) == int(frames):
            traffic, traffic_src = float(ent['traffic_bytes_per_launch']), ent.get('capture')
    except Exception:
        pass
    if sweep:
        for s in sweep:
            s['chain_frac_of_hbm_peak'] = chain.baseline_bytes * s['frames_per_s'] / max(world, 1) / 1e9 / peak
    line = {
        'metric': METRIC[chain.kind], 'value': value, 'unit': 'frames/s',
        'msamples_per_s': value * N / 1e6, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'complex64 (fp32)', 'data': 'synthetic',
        'config': {'workload': w['desc'], 'K': K, 'M': M, 'L': L, 'frames_per_gpu': frames,
                   'constellation': 'QPSK' if chain.kind == 'est_sic' else '16-QAM', 'tx_taps': 'RRC alpha=%g' % w['alpha'],
                   'rx_taps': 'matched filter' if chain.kind == 'est_sic' else RX_TAPS_LABEL,
                   'l2_policy': 'inputs larger than L2 (%.0f MB per buffer vs 126 MB L2)' % (frames * N * 8 / 1e6),
                   'kernels': kernel_names},
        'roofline': {'bound': 'hbm', 'kernel': dom_name + ':' + kname,
                     'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                     'traffic_source': traffic_src, 'peak_source': peak_src, 'algorithmic_bytes_per_launch': alg_bytes,
                     'kernel_ms': {n: ms for (n, _), ms in zip(chain.stages, stage_ms)},
                     'stage_frac': {n: b * frames / (ms * 1e-3) / 1e9 / peak for (n, b), ms in zip(chain.stages, stage_ms)},
                     'chain_algorithmic_bytes_per_frame': chain.baseline_bytes,
                     'chain_achieved_gbs': chain_gbs, 'chain_frac': chain_gbs / peak},
        'clocks': clocks, 'gpu_launches': int(launches),
    }
    if e2e is not None:
        line['e2e'] = e2e
        line['e2e_variants'] = e2e_variants
        line['link_ceiling'] = link
        line['config']['host_affinity'] = affinity_desc
    if latency is not None:
        line['latency'] = latency
    if sweep is not None:
        line['sweep'] = sweep
        line['config']['sweep'] = 'total frames 2^10..2^%d split contiguously over %d rank(s); at most %d frames resident per ' \
                                  'buffer, larger shards loop over the resident buffers' % (args.sweep_max_log2, world, frames)
    if not args.no_cpu:
        ref = cpu_reference_run(w, seconds_budget=args.cpu_seconds / 2.0)
        ref['run']()  # warm
        tcpu = ref['run']()
        v = ref['cores'] * ref['frames_per_thread'] / tcpu
        line['cpu_baseline'] = {'value': v, 'unit': 'frames/s', 'cores': ref['cores'], 'kind': ref['kind'], 'sample': ref['sample']}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
