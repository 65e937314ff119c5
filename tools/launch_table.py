#!/usr/bin/env python
"""Print the last full step of an `ncu --csv` launch list (gpu__time_duration + optional tensor-pipe metrics)."""
import csv
import sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
d = {}
for row in csv.DictReader(lines):
    d.setdefault(int(row['ID']), {'name': row['Kernel Name'].split('(')[0].replace('void ', '')[:24], 'grid': row['Grid Size']})[row['Metric Name']] = float(row['Metric Value'].replace(',', ''))
ids = sorted(d)
starts = [i for i in ids if d[i]['name'].startswith('frontend_mel')]
lo = starts[-2] if len(starts) > 1 and max(ids) - starts[-1] < 40 else starts[-1]
hi = starts[starts.index(lo) + 1] if starts.index(lo) + 1 < len(starts) else max(ids) + 1
tot = 0
for i in ids:
    if lo <= i < hi:
        v = d[i]
        t = v['gpu__time_duration.sum'] / 1000
        tot += t
        print(f"{i - lo:3d} {v['name']:24s} {v['grid']:>14s} {t:8.1f} us  tc {v.get('sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active', 0):5.1f}%  tensor {v.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 0):5.1f}%")
print(f"total {tot:.1f} us")
