#!/usr/bin/env python3
"""bench.py -- GFDM modulate + demodulate throughput on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU kernels (oracle/_ref)

Workload (BASELINE.json configs[2], the configuration the metric is quoted on):
K=1024 subcarriers, M=15 subsymbols, L=2, RRC alpha=0.5 transmit taps, "ZF" receive
taps, random 16-QAM symbols, 4096 frames per GPU.  One *step* = one pass of the hot
path over the batch: modulator_kernel_cc::generic_work on every frame, then
receiver_kernel_cc::generic_work on every modulated frame.

One process per GPU (torchrun for N > 1); frames are independent, so ranks shard the
batch and never communicate on the data path ("scaling": "weak").  Timing: CUDA
events on the launching stream, barrier + synchronize on both sides, max over ranks.
Prints ONE JSON line on rank 0.
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
    # name: (K, M, L, frames per GPU, constellation)
    'c3': dict(K=1024, M=15, L=2, frames=4096, alpha=0.5, seed=1003,
               desc='K=1024 M=15 L=2 16-QAM modulate->ZF-demodulate, 4096 frames/GPU (BASELINE configs[2])'),
    'c1': dict(K=16, M=5, L=2, frames=1 << 18, alpha=0.5, seed=1001,
               desc='K=16 M=5 L=2 modulate->demodulate (BASELINE configs[0])'),
    'c2': dict(K=64, M=9, L=2, frames=1 << 16, alpha=0.2, seed=1002,
               desc='K=64 M=9 L=2 modulate->demodulate (BASELINE configs[1] shape)'),
    'c4': dict(K=256, M=15, L=2, frames=1 << 14, alpha=0.5, seed=1004,
               desc='K=256 M=15 L=2 modulate->demodulate (BASELINE configs[3] shape)'),
    'c5': dict(K=2048, M=15, L=2, frames=2048, alpha=0.5, seed=1005,
               desc='K=2048 M=15 L=2 modulate->demodulate (BASELINE configs[4] shape)'),
}


def dist_env():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


def make_taps(w):
    from gfdm_b200 import design
    tx = design.get_frequency_domain_filter('rrc', w['alpha'], w['M'], w['K'], w['L'])
    rx = design.get_zero_forcing_taps('rrc', w['alpha'], w['M'], w['K'], w['L'])
    return tx.astype(np.complex64), rx.astype(np.complex64)


def make_symbols(w, n_frames, rank):
    from gfdm_b200 import design
    rng = np.random.default_rng(w['seed'] + rank)
    N = w['K'] * w['M']
    # 16-QAM drawn blockwise to bound host memory
    out = np.empty((n_frames, N), np.complex64)
    pts = design.qam16_points().astype(np.complex64)
    step = max(1, (1 << 24) // N)
    for f0 in range(0, n_frames, step):
        f1 = min(n_frames, f0 + step)
        out[f0:f1] = pts[rng.integers(0, 16, (f1 - f0, N))]
    return out


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
    """Time the reference's own CPU kernels (oracle/_ref = unmodified gr-gfdm sources + scalar
    FFTW/VOLK shims; falls back to the plain-C port) on the host cores, one kernel instance per core."""
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
    tx, rx = make_taps(w)
    N = w['K'] * w['M']
    # calibrate on one thread, then size the sample so all threads run ~seconds_budget
    mod = capi.Modulator(w['M'], w['K'], w['L'], tx, lib=lib)
    dem = capi.Demodulator(w['M'], w['K'], w['L'], rx, lib=lib)
    probe = make_symbols(w, 2, 0)
    t0 = time.perf_counter()
    dem.demodulate_batch(mod.modulate_batch(probe))
    per_frame = (time.perf_counter() - t0) / 2
    fpt = int(max(2, min(4096, seconds_budget / max(per_frame, 1e-7))))
    data = make_symbols(w, fpt, 0)
    handles = [(capi.Modulator(w['M'], w['K'], w['L'], tx, lib=lib), capi.Demodulator(w['M'], w['K'], w['L'], rx, lib=lib))
               for _ in range(cores)]
    bufs = [(np.empty_like(data), np.empty_like(data)) for _ in range(cores)]

    def work(i):
        m, d = handles[i]
        x, y = bufs[i]
        m.modulate_batch_host_ptr(x.ctypes.data, data.ctypes.data, fpt)   # ctypes releases the GIL
        d.demodulate_batch_host_ptr(y.ctypes.data, x.ctypes.data, 0, fpt)

    def run_all():
        th = [threading.Thread(target=work, args=(i,)) for i in range(cores)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        return time.perf_counter() - t0

    return dict(kind=kind, cores=cores, frames_per_thread=fpt, run=run_all, N=N,
                sample='%d threads x %d frames of the same workload (mod+demod), %s kernels%s' % (
                    cores, fpt, 'reference gr-gfdm' if kind == 'reference' else 'oracle port',
                    ' + scalar FFTW/VOLK shims' if kind == 'reference' else ''))


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
        'impl': 'reference', 'metric': 'GFDM frames/s (mod+demod)', 'value': value, 'unit': 'frames/s',
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='c3', choices=sorted(WORKLOADS))
    ap.add_argument('--frames', type=int, default=0, help='frames per GPU (default: workload value)')
    ap.add_argument('--cpu-seconds', type=float, default=12.0, help='CPU baseline budget (seconds of wall clock)')
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    ap.add_argument('--no-e2e', action='store_true')
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.frames:
        w['frames'] = args.frames
    if args.impl == 'reference':
        run_reference(args, w)
        return

    import torch
    import torch.distributed as dist
    from gfdm_b200 import capi

    rank, local_rank, world = dist_env()
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    lib = capi.load()
    lib.set_device(local_rank)

    K, M, L, frames = w['K'], w['M'], w['L'], w['frames']
    N = K * M
    tx, rx = make_taps(w)
    mod = capi.Modulator(M, K, L, tx, lib=lib)
    dem = capi.Demodulator(M, K, L, rx, lib=lib)
    stream = torch.cuda.Stream(device=dev)
    mod.set_stream(stream.cuda_stream)
    dem.set_stream(stream.cuda_stream)

    prev_affinity, affinity_desc = bind_to_gpu_numa_node(torch, local_rank)
    host_in = torch.from_numpy(make_symbols(w, frames, rank)).pin_memory()
    d_in = host_in.to(dev, non_blocking=False)
    d_tx = torch.empty_like(d_in)
    d_out = torch.empty_like(d_in)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(events=None):
        if events is not None:
            events[0].record(stream)
        mod.modulate_ptr(d_tx.data_ptr(), d_in.data_ptr(), frames)
        if events is not None:
            events[1].record(stream)
        dem.demodulate_ptr(d_out.data_ptr(), d_tx.data_ptr(), 0, frames)
        if events is not None:
            events[2].record(stream)

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
    launches0 = mod.launch_count() + dem.launch_count()
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        t_start.record(stream)
        for i in range(args.steps):
            step(ev[i])
        t_end.record(stream)
    barrier()
    launches1 = mod.launch_count() + dem.launch_count()
    kernel_names = {'modulator': mod.last_kernel(), 'receiver': dem.last_kernel()}
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
    total_ms = t_start.elapsed_time(t_end)
    mod_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in ev]))
    dem_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in ev]))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    value = world * frames / (ms_per_step * 1e-3)

    # quick sanity of the timed outputs (not a parity test): MF-like recovery up to residual interference
    chk = d_out[:2].cpu().numpy()
    assert np.isfinite(chk).all(), 'non-finite demodulator output'

    # ---- end to end through the C ABI with HOST buffers (H2D + D2H inside the timed region) -------
    e2e = None
    if not args.no_e2e:
        host_tx = torch.empty_like(host_in).pin_memory()
        host_out = torch.empty_like(host_in).pin_memory()

        def e2e_step():
            mod.modulate_batch_host_ptr(host_tx.data_ptr(), host_in.data_ptr(), frames)
            dem.demodulate_batch_host_ptr(host_out.data_ptr(), host_tx.data_ptr(), 0, frames)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / args.e2e_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        bytes_frame = N * 8
        e2e = {'value': world * frames / float(dt.item()), 'unit': 'frames/s',
               'h2d_bytes_per_step': 2 * frames * bytes_frame, 'd2h_bytes_per_step': 2 * frames * bytes_frame,
               'steps': args.e2e_steps, 'ms_per_step': float(dt.item()) * 1e3,
               'path': 'gfdm_modulator_work_batch + gfdm_receiver_work_batch, GFDM_MEM_HOST, pinned host buffers'}

    # ---- the same chain through the byte-wide entries: chunks -> samples -> hard decisions.
    # Reported beside the headline, never instead of it: the symbol side crosses HBM / PCIe as 1 byte per symbol.
    chunk_chain = None
    if not args.no_e2e and N % 16 == 0:
        from gfdm_b200 import design
        sm = capi.Symbol_mapper((design.qam16_points(), capi.DECISION_NEAREST), lib=lib)
        g = torch.Generator(device=dev).manual_seed(w['seed'] + rank)
        d_ch = torch.randint(0, 16, (frames, N), dtype=torch.uint8, device=dev, generator=g)
        d_dec = torch.empty_like(d_ch)
        host_ch = d_ch.cpu().pin_memory()
        host_dec = torch.empty_like(host_ch).pin_memory()
        cev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

        def chunk_step(record=False):
            if record:
                cev[0].record(stream)
            mod.modulate_chunks_ptr(sm, d_tx.data_ptr(), d_ch.data_ptr(), frames)
            if record:
                cev[1].record(stream)
            dem.demodulate_decide_ptr(sm, d_dec.data_ptr(), d_tx.data_ptr(), 0, frames)
            if record:
                cev[2].record(stream)

        with torch.cuda.stream(stream):
            for _ in range(3):
                chunk_step()
            chunk_step(True)
        barrier()
        cm, cd = cev[0].elapsed_time(cev[1]), cev[1].elapsed_time(cev[2])
        chunk_kernels = {'modulator': mod.last_kernel(), 'receiver': dem.last_kernel()}
        # consistency at full size: the fused decisions equal the decisions of the soft-symbol path (bit-exact)
        d_dec2 = torch.empty_like(d_dec)
        dem.demodulate_ptr(d_out.data_ptr(), d_tx.data_ptr(), 0, frames)
        dem.sync()
        sm.decide_ptr(d_dec2.data_ptr(), d_out.data_ptr(), frames * N)
        sm.sync()
        mismatches = int((d_dec != d_dec2).sum().item())

        def chunk_e2e():
            mod.modulate_chunks_batch_host_ptr(sm, host_tx.data_ptr(), host_ch.data_ptr(), frames)
            dem.demodulate_decide_batch_host_ptr(sm, host_dec.data_ptr(), host_tx.data_ptr(), 0, frames)

        chunk_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            chunk_e2e()
        barrier()
        dtc = torch.tensor([(time.perf_counter() - t0) / args.e2e_steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dtc, op=dist.ReduceOp.MAX)
        chunk_chain = {'value': frames / ((cm + cd) * 1e-3), 'unit': 'frames/s per GPU (device resident)',
                       'kernel_ms': {'modulator': cm, 'receiver': cd},
                       'kernels': chunk_kernels,
                       'algorithmic_bytes_per_frame': 2 * (8 * N + N),
                       'achieved_gbs': 2 * 9.0 * N * frames / ((cm + cd) * 1e-3) / 1e9,
                       'decision_mismatches_vs_soft_path': mismatches,
                       'e2e': {'value': world * frames / float(dtc.item()), 'unit': 'frames/s',
                               'h2d_bytes_per_step': frames * 9 * N, 'd2h_bytes_per_step': frames * 9 * N,
                               'ms_per_step': float(dtc.item()) * 1e3,
                               'path': 'gfdm_modulator_work_chunks_batch + gfdm_receiver_work_decide_batch, GFDM_MEM_HOST'}}
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
    alg_bytes = 16.0 * N * frames  # per launch: read N + write N complex64 per frame
    dom, dom_ms = ('modulator', mod_ms) if mod_ms >= dem_ms else ('receiver', dem_ms)
    dom_name = kernel_names[dom]
    # measured DRAM bytes per launch of that kernel from the committed `ncu --set full` capture
    # (tools/ncu_traffic.py -> profiles/ncu_traffic.json); only valid for the frame count it was captured at
    traffic, traffic_src = None, None
    try:
        tdb = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        ent = tdb.get(dom_name)
        if ent and int(ent['frames']) == int(frames):
            traffic, traffic_src = float(ent['traffic_bytes_per_launch']), ent.get('capture')
    except Exception:
        pass
    achieved = alg_bytes / (dom_ms * 1e-3) / 1e9
    chain_gbs = 2 * alg_bytes / ((mod_ms + dem_ms) * 1e-3) / 1e9
    line = {
        'metric': 'GFDM frames/s (mod+demod)', 'value': value, 'unit': 'frames/s',
        'msamples_per_s': value * N / 1e6, 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'complex64 (fp32)', 'data': 'synthetic',
        'config': {'workload': w['desc'], 'K': K, 'M': M, 'L': L, 'frames_per_gpu': frames,
                   'constellation': '16-QAM', 'tx_taps': 'RRC alpha=%g' % w['alpha'], 'rx_taps': 'ZF (Gabor dual, folded to L=2)',
                   'l2_policy': 'inputs larger than L2 (%.0f MB per buffer vs 126 MB L2)' % (frames * N * 8 / 1e6),
                   'kernels': kernel_names},
        'roofline': {'bound': 'hbm', 'kernel': dom + ':' + dom_name,
                     'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                     'traffic_source': traffic_src, 'peak_source': peak_src, 'algorithmic_bytes_per_launch': alg_bytes,
                     'kernel_ms': {'modulator': mod_ms, 'receiver': dem_ms},
                     'chain_achieved_gbs': chain_gbs, 'chain_frac': chain_gbs / peak},
        'clocks': clocks, 'gpu_launches': int(launches),
    }
    if e2e is not None:
        line['e2e'] = e2e
        line['config']['host_affinity'] = affinity_desc
    if chunk_chain is not None:
        line['chunk_chain'] = chunk_chain
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
