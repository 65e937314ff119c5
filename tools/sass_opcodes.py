#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the built library (cuobjdump -sass), restricted to the opcodes that prove the
Blackwell-native path: UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor copies),
UBLKCP (cp.async.bulk), UTCBAR (tcgen05.commit), SYNCS (mbarrier), FFMA2 / FADD2 / FMUL2 (packed fp32), HMMA (mma.sync, none
expected), plus the instruction total.   Usage: sass_opcodes.py lib.so > profiles/r02_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

WANT = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "HMMA",
        "FFMA", "LDS", "STS", "LDG", "STG", "BAR", "SHFL", "REDG", "ATOMG"]
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        hist[kern]["total"] += 1
        for w in WANT:
            if op == w or (w not in ("FFMA", "BAR") and op.startswith(w)):
                hist[kern][w] += 1
                break
cols = ["total"] + WANT
print(f"{'kernel':58s} " + " ".join(f"{c:>7s}" for c in cols))
for k, h in hist.items():
    print(f"{k[:58]:58s} " + " ".join(f"{h.get(c, 0):7d}" for c in cols))
tot = collections.Counter()
for h in hist.values():
    tot.update(h)
print(f"{'ALL KERNELS':58s} " + " ".join(f"{tot.get(c, 0):7d}" for c in cols))
