"""Time the VII and generic-grid geo kernels on device-resident data:  python tools/sibling_probe.py [granules]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-geotiepoints_b200"))
import numpy as np, torch
from geotiepoints_b200 import testing, viiinterpolator
from geotiepoints_b200.geointerpolator import GeoGridInterpolator
g = int(sys.argv[1]) if len(sys.argv) > 1 else 4
parts = [testing.modis_granule(k) for k in range(g)]
lon = torch.from_numpy(np.concatenate([p[0] for p in parts])).cuda()
lat = torch.from_numpy(np.concatenate([p[1] for p in parts])).cuda()
n = lon.shape[0]
grid = GeoGridInterpolator((4.0 * np.arange(n) + 1.5, 4.0 * np.arange(1354)), lon, lat, bounds_error=False, fill_value=None)
fine = (np.arange(4.0 * n), np.arange(4.0 * 1354))
cases = {"vii_geo": lambda: viiinterpolator.tie_points_geo_interpolation(lon, lat, 10, 4, lat_threshold_use_cartesian=-1.0),
         "geogrid_linear": lambda: grid.interpolate(fine),
         "geogrid_cubic": lambda: grid.interpolate(fine, method="cubic")}
for name, fn in cases.items():
    for _ in range(2):
        out = fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(name, f"{dt * 1e3:.2f} ms", f"{out[0].numel() / dt:.3g} px/s")
