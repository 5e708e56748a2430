#!/usr/bin/env python3
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the per-kernel lines the roofline discussion uses.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; tools/ncu_summary.py raw.csv > profiles/<name>.txt"""
import csv
import sys

KEYS = [
    ('gpu__time_duration.sum', 'duration'),
    ('dram__bytes_read.sum', 'dram read'),
    ('dram__bytes_write.sum', 'dram write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram throughput % of peak'),
    ('lts__t_bytes.sum', 'L2 bytes'),
    ('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smem wavefronts'),
    ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts'),
    ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex throughput %'),
    ('smsp__inst_executed.sum', 'warp instructions'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
    ('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'fma pipe %'),
    ('sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'lsu pipe %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'achieved occupancy %'),
    ('launch__registers_per_thread', 'registers/thread'),
    ('launch__shared_mem_per_block_dynamic', 'dynamic smem/block'),
    ('launch__grid_size', 'grid'),
    ('launch__block_size', 'block'),
    ('sm__cycles_elapsed.max', 'sm cycles elapsed'),
]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('kernel: %s' % d.get('Kernel Name', '?'))
    for k, label in KEYS:
        if k in d:
            print('  %-28s %s %s   [%s]' % (label, d[k], units[hdr.index(k)], k))
    stalls = []
    for k in hdr:
        if k.startswith('smsp__average_warps_issue_stalled_') and k.endswith('_per_issue_active.ratio'):
            try:
                stalls.append((float(d[k]), k[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    print('  stalls (warps per issue-active cycle): ' + ', '.join('%s %.2f' % (n, v) for v, n in stalls[:8]))
    try:
        rd = float(d['dram__bytes_read.sum']); wr = float(d['dram__bytes_write.sum'])
        scale = {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1.0}
        rd *= scale[units[hdr.index('dram__bytes_read.sum')]]; wr *= scale[units[hdr.index('dram__bytes_write.sum')]]
        t = float(d['gpu__time_duration.sum']) * {'us': 1e-6, 'ms': 1e-3, 'ns': 1e-9, 's': 1.0}[units[hdr.index('gpu__time_duration.sum')]]
        print('  traffic = %.1f MB per launch -> %.0f GB/s under the profiler' % ((rd + wr) / 1e6, (rd + wr) / t / 1e9))
    except Exception:
        pass
    print()
