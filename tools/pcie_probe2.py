#!/usr/bin/env python
"""What bounds the end-to-end (numpy in / numpy out) path at N GPUs of one box: the host side of PCIe.

    python tools/pcie_probe2.py [n_gpus]        (gpurun --gpus 8; one process per GPU, all copying at once)

For every experiment each GPU moves a 1 GiB buffer `reps` times between device memory and page-locked host
memory, all processes starting at the same wall-clock instant; the table gives per-GPU and summed GB/s.
Experiments: D2H with 1, 2, 4 and all GPUs (default page-locked memory); D2H into write-combined page-locked
memory; H2D; both directions at once; D2H with the copies cut into 64 MiB pieces on 3 streams (what the host
entry points of libgtp_b200 do); and, for context, what the host cores themselves can write (numpy copy, one
thread per GPU).  Also prints the topology nvidia-smi reports and the NUMA layout of the host.
"""
import os
import subprocess
import sys
import time

GIB = 1 << 30


def worker(idx, plan_path):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "python-geotiepoints_b200"))
    import json
    import numpy as np
    from cuda.bindings import runtime as rt
    from geotiepoints_b200 import _capi
    _capi.bind_host_thread_to_gpu(idx)
    plan = json.load(open(plan_path))

    def ok(res):
        err, *rest = res
        if int(err) != 0:
            raise RuntimeError(f"CUDA error {err}")
        return rest[0] if len(rest) == 1 else rest

    ok(rt.cudaSetDevice(idx))
    dev = ok(rt.cudaMalloc(GIB))
    dev2 = ok(rt.cudaMalloc(GIB))
    host = ok(rt.cudaHostAlloc(GIB, rt.cudaHostAllocPortable))
    host2 = ok(rt.cudaHostAlloc(GIB, rt.cudaHostAllocPortable))
    host_wc = ok(rt.cudaHostAlloc(GIB, rt.cudaHostAllocPortable | rt.cudaHostAllocWriteCombined))
    streams = [ok(rt.cudaStreamCreate()) for _ in range(3)]
    D2H, H2D = rt.cudaMemcpyKind.cudaMemcpyDeviceToHost, rt.cudaMemcpyKind.cudaMemcpyHostToDevice
    ok(rt.cudaMemcpyAsync(host, dev, GIB, D2H, streams[0]))
    ok(rt.cudaMemcpyAsync(host_wc, dev, GIB, D2H, streams[0]))
    ok(rt.cudaMemcpyAsync(dev2, host2, GIB, H2D, streams[0]))
    ok(rt.cudaDeviceSynchronize())
    a = np.ones(GIB // 16)            # 512 MiB
    b = np.empty_like(a)
    for name, t_start, members, reps in plan:
        if idx not in members:
            continue
        while time.time() < t_start:
            time.sleep(0.0005)
        t0 = time.perf_counter()
        moved = reps * GIB
        if name.startswith("d2h_wc"):
            for _ in range(reps):
                ok(rt.cudaMemcpyAsync(host_wc, dev, GIB, D2H, streams[0]))
        elif name.startswith("d2h_chunked"):
            piece = 64 << 20
            for _ in range(reps):
                for k in range(GIB // piece):
                    ok(rt.cudaMemcpyAsync(host + k * piece, dev + k * piece, piece, D2H, streams[k % 3]))
        elif name.startswith("d2h"):
            for _ in range(reps):
                ok(rt.cudaMemcpyAsync(host, dev, GIB, D2H, streams[0]))
        elif name.startswith("h2d"):
            for _ in range(reps):
                ok(rt.cudaMemcpyAsync(dev2, host2, GIB, H2D, streams[0]))
        elif name.startswith("bidir"):
            for _ in range(reps):
                ok(rt.cudaMemcpyAsync(host, dev, GIB, D2H, streams[0]))
                ok(rt.cudaMemcpyAsync(dev2, host2, GIB, H2D, streams[1]))
            moved *= 2
        elif name.startswith("cpu_copy"):
            for _ in range(reps):
                np.copyto(b, a)
            moved = reps * a.nbytes
        ok(rt.cudaDeviceSynchronize())
        dt = time.perf_counter() - t0
        print(f"RESULT {name} {idx} {moved / dt / 1e9:.2f}", flush=True)


def main():
    import json
    import tempfile
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"], ["numactl", "--hardware"]):
        try:
            out = subprocess.run(cmd, capture_output=True, text=True, timeout=30).stdout
            if cmd[0] == "lscpu":
                out = "\n".join(line for line in out.splitlines() if any(k in line for k in ("Model name", "Socket", "NUMA", "CPU(s):", "Thread")))
            print(f"$ {' '.join(cmd)}\n{out}")
        except Exception as exc:
            print(f"$ {' '.join(cmd)}: {exc!r}")
    t = time.time() + 35.0
    plan = []

    def add(name, members, reps=6, gap=6.0):
        nonlocal t
        plan.append((name, t, list(members), reps))
        t += gap
    for k in (1, 2, 4, n):
        if k <= n:
            add(f"d2h_{k}gpus", range(k))
    if n >= 8:
        add("d2h_gpus0-3", range(0, 4))
        add("d2h_gpus4-7", range(4, 8))
        add("d2h_even_gpus", range(0, 8, 2))
    add(f"d2h_wc_{n}gpus", range(n))
    add(f"d2h_chunked_{n}gpus", range(n))
    add(f"h2d_{n}gpus", range(n))
    add(f"bidir_{n}gpus", range(n))
    add(f"cpu_copy_{n}threads", range(n), reps=3, gap=10.0)
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as fh:
        json.dump(plan, fh)
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", str(i), fh.name],
                              stdout=subprocess.PIPE, text=True) for i in range(n)]
    results = {}
    for p in procs:
        out, _ = p.communicate()
        for line in out.splitlines():
            if line.startswith("RESULT"):
                _, name, idx, gbs = line.split()
                results.setdefault(name, {})[int(idx)] = float(gbs)
    print(f"{'experiment':24s} {'sum GB/s':>9s}   per GPU")
    for name, _, members, _ in plan:
        r = results.get(name, {})
        print(f"{name:24s} {sum(r.values()):9.1f}   " + " ".join(f"{i}:{r.get(i, float('nan')):.1f}" for i in members))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--worker":
        worker(int(sys.argv[2]), sys.argv[3])
    else:
        main()
