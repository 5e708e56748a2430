#!/usr/bin/env python3
"""Extract per-launch DRAM traffic of the fused kernels from an `ncu --set full` capture.
usage: ncu -i X.ncu-rep --page raw --csv > raw.csv; tools/ncu_traffic.py raw.csv <frames> <source-tag> >> merged into
profiles/ncu_traffic.json (read by bench.py for roofline.traffic; only used when the frame count matches)."""
import csv
import json
import os
import re
import sys

raw, frames, tag = sys.argv[1], int(sys.argv[2]), sys.argv[3]
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'ncu_traffic.json')
db = json.load(open(path)) if os.path.exists(path) else {}
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    m = re.match(r'void (fused_\w+)<Shape<(\d+), (\d+), (\d+), (\d+)', d['Kernel Name'])
    if not m:
        m2 = re.match(r'void (fused_\w+)<Shape2<(\d+), (\d+), (\d+), (\d+), (\d+)', d['Kernel Name'])
        if not m2:
            continue
        name = '%s<M=%s,K=%sx%sx%s,T=%s>' % (m2.group(1), m2.group(2), m2.group(3), m2.group(4), m2.group(5), m2.group(6))
    else:
        name = '%s<M=%s,K=%sx%s,T=%s>' % m.groups()
    rd = float(d['dram__bytes_read.sum']) * UNIT[u['dram__bytes_read.sum']]
    wr = float(d['dram__bytes_write.sum']) * UNIT[u['dram__bytes_write.sum']]
    db[name] = {'traffic_bytes_per_launch': rd + wr, 'dram_read': rd, 'dram_write': wr, 'frames': frames,
                'duration_us_under_ncu': float(d['gpu__time_duration.sum']), 'capture': tag}
json.dump(db, open(path, 'w'), indent=1, sort_keys=True)
print(json.dumps(db, indent=1, sort_keys=True))
