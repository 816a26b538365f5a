#!/usr/bin/env python
"""Per-source-line executed-instruction and stall-sample shares of an .ncu-rep captured with
--import-source on (kernels built with -lineinfo).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [min_percent]
"""
import collections
import csv
import io
import subprocess
import sys


def main(rep, min_pct=0.4):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur, ln, src = None, None, ""
    data = collections.defaultdict(lambda: collections.defaultdict(lambda: [0, 0, "", collections.Counter()]))
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] in ("Function Name", "Kernel Name", "Line No"):
            continue
        if r[0].strip().isdigit():
            ln, src = int(r[0]), r[1]
            data[cur][ln][2] = src
            continue
        if r[0] == "" and len(r) > 7 and r[2].startswith("0x"):
            d = data[cur][ln]
            d[0] += int(r[7] or 0)
            d[1] += int(r[6] or 0)
            parts = r[3].split()
            op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
            d[3][op.split(".")[0]] += int(r[7] or 0)
    tot = sum(d[0] for v in data.values() for d in v.values())
    smp = sum(d[1] for v in data.values() for d in v.values())
    print(f"total warp instructions {tot}, stall samples {smp}")
    for f, v in data.items():
        ft = sum(d[0] for d in v.values())
        print(f"== {f}: {100 * ft / tot:.1f} % of instructions, {100 * sum(d[1] for d in v.values()) / max(smp, 1):.1f} % of samples")
        for ln, d in sorted(v.items()):
            if d[0] > min_pct / 100 * tot:
                ops = " ".join(f"{o}:{100 * n / tot:.1f}" for o, n in d[3].most_common(3))
                print(f"  {ln:4d} {100 * d[0] / tot:5.2f}% inst {100 * d[1] / max(smp, 1):5.2f}% smp | {d[2].strip()[:90]} [{ops}]")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.4)
