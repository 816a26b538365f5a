"""Time the three legacy kernels on device-resident data:  python tools/legacy_probe.py [granules]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-geotiepoints_b200"))
import numpy as np, torch
from geotiepoints_b200 import legacy, testing
g = int(sys.argv[1]) if len(sys.argv) > 1 else 4
parts = [testing.modis_granule(k) for k in range(g)]
lon = torch.from_numpy(np.concatenate([p[0] for p in parts])).cuda()
lat = torch.from_numpy(np.concatenate([p[1] for p in parts])).cuda()
for fn in (legacy.modis1kmto250m, legacy.modis1kmto500m):
    for _ in range(2):
        out = fn(lon, lat)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        out = fn(lon, lat)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(fn.__name__, f"{dt * 1e3:.2f} ms", f"{out[0].numel() / dt:.3g} px/s")
