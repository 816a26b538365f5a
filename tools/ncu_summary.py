#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into the few numbers the roofline argument needs.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.per_cycle_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__inst_executed.sum", "smsp__inst_executed.sum", "thread_inst_executed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    rows = ncu(rep, "raw")
    hdr, units = rows[0], rows[1]
    for k, vals in enumerate(rows[2:]):
        d = dict(zip(hdr, zip(units, vals)))
        print(f"== launch {k}: {d.get('Kernel Name', ('', '?'))[1]}")
        for key in KEYS:
            if key in d:
                print(f"{key:75s} {d[key][1]:>18s} {d[key][0]}")
        for key in sorted(d):
            if key.startswith("smsp__average_warps_issue_stalled") and key.endswith("_per_issue_active.ratio"):
                v = float(d[key][1] or 0)
                if v >= 0.05:
                    print(f"stall {key[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:30s} {v:8.3f} warps per issue")
    src = ncu(rep, "source")
    if len(src) > 2:
        hdr = src[1]
        ix = {h: i for i, h in enumerate(hdr)}
        hist, tot = collections.Counter(), 0
        for r in src[2:]:
            if len(r) <= ix["Instructions Executed"]:
                continue
            parts = r[ix["Source"]].strip().split()
            if not parts:
                continue
            op = parts[1] if parts[0].startswith("@") and len(parts) > 1 else parts[0]
            n = int(r[ix["Instructions Executed"]] or 0)
            hist[op.split(".")[0]] += n
            tot += n
        print("== executed warp instructions by opcode (first launch in the report)")
        for op, n in hist.most_common(16):
            print(f"{op:10s} {n:14d} {100.0 * n / max(tot, 1):5.1f} %")


if __name__ == "__main__":
    main(sys.argv[1])
