#!/usr/bin/env python
"""Pinned D2H rate of every GPU of the box, all copying at the same time (what bounds the end-to-end
numpy path at N GPUs):  python tools/pcie_concurrent.py [n_gpus]   (spawns one process per GPU)"""
import os
import subprocess
import sys
import time

if len(sys.argv) > 1 and sys.argv[1] == "--worker":
    idx, t_start = int(sys.argv[2]), float(sys.argv[3])
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "python-geotiepoints_b200"))
    import torch
    from geotiepoints_b200 import _capi
    bound = _capi.bind_host_thread_to_gpu(idx)
    dev = torch.device("cuda", idx)
    n = 1 << 30
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.copy_(d, non_blocking=True); torch.cuda.synchronize(dev)
    while time.time() < t_start:      # all workers start together
        time.sleep(0.001)
    t0 = time.perf_counter()
    reps = 6
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    print(f"gpu {idx}: {reps * n / dt / 1e9:6.1f} GB/s D2H  (cpus {bound[0] if bound else None}..{bound[-1] if bound else None})", flush=True)
    sys.exit(0)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
t_start = time.time() + 25.0
procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", str(i), str(t_start)]) for i in range(n)]
for p in procs:
    p.wait()
