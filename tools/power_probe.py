#!/usr/bin/env python
"""Sustained SM clock and board power of single-resource loads (FP64 FMA pipe, FP32 FMA pipe,
HBM writes), read through NVML while tools/microbench keeps the resource busy.  Used to split
the power budget of the interpolation kernel ("power")."""
import subprocess
import sys
import time

import pynvml

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
for op in sys.argv[1:] or ["dfma", "ffma", "write"]:
    proc = subprocess.Popen(["./tools/microbench", "loop", op, "2.5"], stdout=subprocess.PIPE, text=True)
    t0, rows = time.time(), []
    while proc.poll() is None:
        rows.append((time.time() - t0, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                     pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.02)
    tail = [r for r in rows if r[0] > 1.5]
    mhz = sorted(r[1] for r in tail)[len(tail) // 2] if tail else None
    watts = sorted(r[2] for r in tail)[len(tail) // 2] if tail else None
    print(f"{op:6s} sustained: {mhz} MHz, {watts:.0f} W | {proc.stdout.read().strip()}")
    time.sleep(1.0)
