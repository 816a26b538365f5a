#!/usr/bin/env python
"""Per-source-line executed-instruction counts of the first kernel in an .ncu-rep (needs -lineinfo
and --import-source on):  python tools/ncu_lines.py report.ncu-rep [units] [min_per_unit]
`units` divides the counts (e.g. the number of quad-row warps) so the numbers read "per quad"."""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
floor = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, agg, samples = None, None, collections.OrderedDict(), {}
ops = collections.Counter()
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Function Name":
        continue
    if r and r[0] == "Line No":
        hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if hdr is None or len(r) < len(hdr) - 2:
        continue
    try:
        c = int(r[iI])
    except ValueError:
        continue
    if r[2] == "-":
        if c:
            agg[(cur, int(r[0]))] = (c, int(r[iS] or 0), r[1][:100])
    else:
        op = re.sub(r"^\s*(@!?U?P\d+\s+)?", "", r[3]).split()
        if op:
            ops[op[0].split(".")[0]] += c
tot = sum(v[0] for v in agg.values())
tots = sum(v[1] for v in agg.values()) or 1
print(f"total {tot / units:.1f} per unit")
for (f, l), (c, smp, src) in agg.items():
    if c / units >= floor:
        print(f"{f:22s}:{l:4d} {c / units:7.1f} {100.0 * smp / tots:5.1f}%  {src}")
print("== opcodes per unit")
for op, c in ops.most_common(30):
    print(f"   {op:10s} {c / units:7.1f}")
