#!/usr/bin/env python
"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum` launch list by kernel name for the LAST step (from the last
frontend_mel launch on): count, total us, share.  With --split the step is cut at the first loss kernel (build_targets) into
forward and backward."""
import csv, sys, collections
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
rows = [r for r in csv.DictReader(lines) if r['Metric Name'] == 'gpu__time_duration.sum']
starts = [int(r['ID']) for r in rows if 'frontend_mel' in r['Kernel Name']]
lo = starts[-1]
rows = [r for r in rows if int(r['ID']) >= lo]
cut = next((int(r['ID']) for r in rows if 'build_targets' in r['Kernel Name']), None)
parts = [("step", rows)]
if '--split' in sys.argv and cut is not None:
    parts = [("forward", [r for r in rows if int(r['ID']) < cut]), ("loss + backward + optimizer", [r for r in rows if int(r['ID']) >= cut])]
for title, rs in parts:
    agg = collections.OrderedDict(); tot = 0.0
    for r in rs:
        name = r['Kernel Name'].split('(')[0].replace('void ', '').replace('yad::', '')[:60]
        t = float(r['Metric Value'].replace(',', '')) / 1000
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t; tot += t
    print(f"== {title}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:62s} x{n:4d} {t:10.1f} us {100*t/tot:5.1f}%")
    print(f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches")
