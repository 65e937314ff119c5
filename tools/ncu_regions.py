#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` dump: per barrier-delimited region of the SASS,
instructions executed, shared-memory wavefronts (ideal / excess) and stall samples.  Usage: ncu_regions.py dump.csv"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
def g(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0
reg, regs = None, []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]]
    if reg is None or "BAR." in src or "SYNCS" in src or "WARPSYNC" in src and False:
        reg = {"start": r[ix["Address"]], "first": src.strip()[:50], "n": 0, "inst": 0, "wf": 0, "wfx": 0, "samp": 0, "bar": 0,
               "ssb": 0, "lsb": 0, "top": []}
        regs.append(reg)
    reg["n"] += 1
    reg["inst"] += g(r, "Instructions Executed")
    reg["wf"] += g(r, "L1 Wavefronts Shared")
    reg["wfx"] += g(r, "L1 Wavefronts Shared Excessive")
    s = g(r, "# Samples")
    reg["samp"] += s
    reg["bar"] += g(r, "stall_barrier")
    reg["ssb"] += g(r, "stall_short_sb")
    reg["lsb"] += g(r, "stall_long_sb")
    reg["top"].append((s, src.strip()[:60], g(r, "L1 Wavefronts Shared Excessive")))
tot = sum(x["samp"] for x in regs) or 1
print(f"{'addr':>8} {'#sass':>5} {'inst':>12} {'smem wf':>11} {'excess':>10} {'samples':>8} {'%':>5} {'barrier':>7} {'short_sb':>8} {'long_sb':>7}  first")
for x in regs:
    print(f"{x['start'][-6:]:>8} {x['n']:5d} {x['inst']:12.0f} {x['wf']:11.0f} {x['wfx']:10.0f} {x['samp']:8.0f} {100 * x['samp'] / tot:5.1f} "
          f"{x['bar']:7.0f} {x['ssb']:8.0f} {x['lsb']:7.0f}  {x['first']}")
if len(sys.argv) > 2:
    for x in regs:
        print("==", x["start"], x["first"])
        for s, src, e in sorted(x["top"], reverse=True)[:int(sys.argv[2])]:
            print(f"   {s:7.0f} {e:9.0f}  {src}")
