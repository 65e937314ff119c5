#!/usr/bin/env python
"""Per-source-line view of an ncu `--set full --import-source on` capture, built offline (no GPU):
joins the SASS page of the report (per-instruction samples / stall reasons / shared-memory wavefronts) with the line table of
the cubin (`nvdisasm -g`) by instruction offset.

    ncu_by_line.py report.ncu-rep lib.so kernel_substring [--min 0.004] > profiles/<name>_by_line.txt
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

WANT = ["# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Ideal", "stall_barrier", "stall_short_sb",
        "stall_long_sb", "stall_wait", "stall_mio", "stall_math", "stall_not_selected", "stall_selected", "stall_dispatch", "stall_lg",
        "stall_branch_resolving", "stall_no_inst", "stall_sleep", "stall_membar"]


def sass_rows(rep, kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            blocks.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] is not None and r and r[0].startswith("0x"):
            cur["rows"].append(r)
    for b in blocks:
        if kernel in b["name"]:
            return b
    raise SystemExit(f"kernel {kernel!r} not in report: {[b['name'][:60] for b in blocks]}")


def line_table(so, kernel_mangled_part):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, capture_output=True)
    table = None
    for f in sorted(os.listdir(d)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(d, f)], capture_output=True, text=True).stdout
        if kernel_mangled_part not in txt:
            continue
        cur_fn, line, tabs = None, None, {}
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                cur_fn, line = m.group(1), None
                tabs[cur_fn] = {}
                continue
            if ln.lstrip().startswith("//## File"):
                # innermost location first, then "inlined at" ...: attribute to the OUTERMOST call site (the kernel body's line)
                locs = re.findall(r'"([^"]+)", line (\d+)', ln)
                if locs:
                    line = (os.path.basename(locs[-1][0]), int(locs[-1][1]))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and cur_fn is not None:
                tabs[cur_fn][int(m.group(1), 16)] = (line, m.group(2).strip())
        for fn, t in tabs.items():
            if kernel_mangled_part in fn:
                table = t if table is None or len(t) > len(table) else table
                return fn, t
    raise SystemExit("kernel not found in any cubin")


def main():
    rep, so, kern = sys.argv[1:4]
    mangled = sys.argv[4] if len(sys.argv) > 4 and not sys.argv[4].startswith("--") else kern
    thr = float(sys.argv[sys.argv.index("--min") + 1]) if "--min" in sys.argv else 0.004
    b = sass_rows(rep, kern)
    ix = {h: i for i, h in enumerate(b["hdr"])}
    fn, tab = line_table(so, mangled)
    base = int(b["rows"][0][0], 16)
    per, tot, order = {}, {w: 0.0 for w in WANT}, []
    for r in b["rows"]:
        off = int(r[0], 16) - base
        line, _ = tab.get(off, (None, ""))
        key = line or ("?", 0)
        if key not in per:
            per[key] = {w: 0.0 for w in WANT}
            per[key]["n_sass"] = 0
            order.append(key)
        per[key]["n_sass"] += 1
        for w in WANT:
            if w in ix:
                try:
                    v = float(r[ix[w]])
                except ValueError:
                    v = 0.0
                per[key][w] += v
                tot[w] += v
    print(f"# {b['name'][:100]}\n# cubin function {fn}; {len(b['rows'])} SASS instructions")
    print("# totals: " + ", ".join(f"{w}={tot[w]:.0f}" for w in WANT if tot[w]))
    src_cache = {}

    def src(key):
        f, n = key
        for root in ("yolo-inspired-audio-activity-detection_b200/csrc", "include"):
            p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), root, f)
            if os.path.exists(p):
                if p not in src_cache:
                    src_cache[p] = open(p).read().splitlines()
                return src_cache[p][n - 1].strip()[:80] if 0 < n <= len(src_cache[p]) else ""
        return ""
    cols = ["# Samples", "Instructions Executed", "L1 Wavefronts Shared", "stall_barrier", "stall_short_sb", "stall_long_sb", "stall_wait",
            "stall_mio", "stall_math", "stall_not_selected", "stall_selected", "stall_dispatch", "stall_branch_resolving", "stall_no_inst"]
    print("%-18s %5s %7s %7s %7s | %5s %5s %5s %5s %5s %5s %5s %5s %5s %5s %5s  source" %
          ("file:line", "sass", "smp%", "inst%", "wavef%", "barr", "ssb", "lsb", "wait", "mio", "math", "nsel", "sel", "disp", "brch", "noin"))
    for key in sorted(order, key=lambda k: (k[0], k[1])):
        v = per[key]
        if v["# Samples"] < thr * tot["# Samples"] and v["Instructions Executed"] < thr * tot["Instructions Executed"] and \
                v["L1 Wavefronts Shared"] < thr * max(1.0, tot["L1 Wavefronts Shared"]):
            continue
        pct = lambda w: 100.0 * v[w] / tot[w] if tot[w] else 0.0   # noqa: E731
        sm = max(1.0, v["# Samples"])
        print("%-18s %5d %7.2f %7.2f %7.2f | " % (f"{key[0]}:{key[1]}", v["n_sass"], pct("# Samples"), pct("Instructions Executed"),
                                                    pct("L1 Wavefronts Shared")) +
              " ".join("%5.0f" % (100.0 * v[c] / sm) for c in cols[3:]) + "  " + src(key))


if __name__ == "__main__":
    main()
