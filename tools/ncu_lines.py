#!/usr/bin/env python3
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line:
stall samples, executed instructions and the dominant stall reasons.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:K > f.csv; ncu_lines.py f.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur = None
hdr = None
agg = []
for r in rows:
    if len(r) >= 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or not r or r[0] == '' or r[0] == 'Function Name':
        continue
    d = dict(zip(hdr, r))
    try:
        s = int(d['# Samples'])
        ins = int(d['Instructions Executed'])
    except (KeyError, ValueError):
        continue
    stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith('stall_') and 'Not Issued' not in k and v.isdigit() and int(v) > 0}
    agg.append((s, ins, cur, d['Line No'], r[1].strip()[:95], stalls))
tot = sum(a[0] for a in agg)
toti = sum(a[1] for a in agg)
print('total samples %d, warp instructions %d' % (tot, toti))
allst = {}
for a in agg:
    for k, v in a[5].items():
        allst[k] = allst.get(k, 0) + v
print('stalls:', ', '.join('%s %.1f%%' % (k, 100.0 * v / max(1, sum(allst.values()))) for k, v in sorted(allst.items(), key=lambda x: -x[1])))
for s, ins, f, l, src, st in sorted(agg, reverse=True)[:top]:
    ss = ' '.join('%s:%d' % (k, v) for k, v in sorted(st.items(), key=lambda x: -x[1])[:3])
    print('%6d %5.1f%% inst %7d  %s:%s  %s   [%s]' % (s, 100.0 * s / max(tot, 1), ins, f, l, src, ss))
