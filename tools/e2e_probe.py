#!/usr/bin/env python
"""Where does the end-to-end (host numpy -> host numpy) time of the public API go?

    python tools/e2e_probe.py            # raw PCIe rates, then the API at several chunk sizes

Prints (i) the raw pinned-memory D2H / H2D / bidirectional copy rates of this box (the ceiling
of any host-to-host path: 16 B of float64 lon+lat leave the GPU per output pixel), (ii) the
public API `modisinterpolator.modis_1km_to_250m(numpy)` timed for 1 and 4 granules per call at
several values of GTP_CHUNK_MB (the pipeline's chunk size, one subprocess each because the
library reads the variable once).
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-geotiepoints_b200"))


def raw_rates():
    import torch
    dev = torch.device("cuda", 0)
    n = 1 << 30
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def timed(fn, reps=5):
        fn(); torch.cuda.synchronize()
        best = 1e9
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return best

    out = {}
    out["d2h_GBs"] = n / timed(lambda: h.copy_(d, non_blocking=True)) / 1e9
    out["h2d_GBs"] = n / timed(lambda: d.copy_(h, non_blocking=True)) / 1e9

    def both():
        with torch.cuda.stream(s1):
            h.copy_(d, non_blocking=True)
        with torch.cuda.stream(s2):
            d2.copy_(h2, non_blocking=True)
    t = timed(both)
    out["bidir_each_GBs"] = n / t / 1e9

    def two_d2h():
        half = n // 2
        with torch.cuda.stream(s1):
            h[:half].copy_(d[:half], non_blocking=True)
        with torch.cuda.stream(s2):
            h[half:].copy_(d[half:], non_blocking=True)
    out["d2h_two_streams_GBs"] = n / timed(two_d2h) / 1e9
    for mb in (4, 16, 64):
        k = mb << 20

        def chunks():
            for o in range(0, n, k):
                h[o:o + k].copy_(d[o:o + k], non_blocking=True)
        out[f"d2h_{mb}MB_chunks_GBs"] = n / timed(chunks, reps=3) / 1e9
    return out


def api_run(granules, reps):
    import numpy as np
    import torch
    from geotiepoints_b200 import _capi, modisinterpolator, testing
    _capi.bind_host_thread_to_gpu(0)
    arrs = testing.modis_granule(3, dtype=np.float64)
    arrs = [np.concatenate([a] * granules) for a in arrs]
    pinned = []
    for a in arrs:
        p = _capi.pinned_pool.empty(a.shape, a.dtype)
        p[...] = a
        pinned.append(p)
    px = pinned[0].shape[0] * 4 * 5416
    res = modisinterpolator.modis_1km_to_250m(*pinned); del res
    res = modisinterpolator.modis_1km_to_250m(*pinned); del res
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        res = modisinterpolator.modis_1km_to_250m(*pinned)
        _ = float(res[0][-1, -1])
        ts.append(time.perf_counter() - t0)
        del res
    # pageable inputs and outputs handed to a fresh (non-pooled) consumer
    t0 = time.perf_counter()
    res = modisinterpolator.modis_1km_to_250m(*arrs)
    t_pageable_in = time.perf_counter() - t0
    del res
    best = min(ts)
    return {"granules": granules, "chunk_mb": os.environ.get("GTP_CHUNK_MB", "default"),
            "best_ms": best * 1e3, "median_ms": sorted(ts)[len(ts) // 2] * 1e3,
            "px_per_s": px / best, "d2h_GBs": 16 * px / best / 1e9,
            "pageable_inputs_ms": t_pageable_in * 1e3}


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--api":
        print(json.dumps(api_run(int(sys.argv[2]), int(sys.argv[3]))), flush=True)
        sys.exit(0)
    print(json.dumps({"raw": raw_rates()}), flush=True)
    for chunk in ("8", "16", "32", "64", "128"):
        for granules in (1, 4):
            env = dict(os.environ, GTP_CHUNK_MB=chunk)
            subprocess.run([sys.executable, os.path.abspath(__file__), "--api", str(granules), "5"], env=env)
