#!/usr/bin/env python
"""Every kernel of libgtp_b200 once, on small inputs - meant to run under compute-sanitizer where it is available
(it is closed on the gpurun pool; tests/test_gpu_parity.py::test_kernels_stay_inside_their_output_buffers
is the stand-in there):
    compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
    compute-sanitizer --tool racecheck python tools/sanitize_smoke.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-geotiepoints_b200"))
import numpy as np
import torch
from geotiepoints_b200 import mod03, modisinterpolator, simple_modis_interpolator, testing

torch.cuda.set_device(0)
n = 0
for granule in (0, 4):          # equatorial; polar (wide / slow / wrap paths)
    lon, lat, satz = testing.modis_granule(granule, n_scans=3)
    if granule == 4:
        lon, lat, satz = (np.ascontiguousarray(a[-30:]) for a in testing.modis_granule(4))
    for dtype in (np.float64, np.float32):
        a = [x.astype(dtype) for x in (lon, lat, satz)]
        dev = [torch.from_numpy(x).cuda() for x in a]
        modisinterpolator.modis_1km_to_250m(*dev); n += 1
        modisinterpolator.modis_1km_to_500m(*dev); n += 1
        simple_modis_interpolator.modis_1km_to_250m(dev[0], dev[1]); n += 1
        simple_modis_interpolator.modis_1km_to_500m(dev[0], dev[1]); n += 1
        for width in (271, 270):
            a5 = testing.subsample_5km(*a, width=width)
            modisinterpolator.modis_5km_to_1km(*[torch.from_numpy(x).cuda() for x in a5]); n += 1
        modisinterpolator.modis_1km_to_250m(*a); n += 1            # host pipeline
    f = [x.astype(np.float32) for x in (lon, lat)]
    zen = np.round(satz / 0.01).astype(np.int16)
    for out_dtype in (np.float32, np.float64):
        mod03.modis_1km_to_250m(*f, zen, zenith_scale=0.01, out_dtype=out_dtype); n += 1
        mod03.modis_1km_to_500m(*f, satz.astype(np.float32), out_dtype=out_dtype); n += 1
        mod03.simple_modis_1km_to_250m(*f, out_dtype=out_dtype); n += 1
        l5, a5, z5 = testing.subsample_5km(f[0], f[1], zen, 270)
        mod03.modis_5km_to_1km(l5, a5, z5, zenith_scale=0.01, out_dtype=out_dtype); n += 1
torch.cuda.synchronize()
print(f"sanitize_smoke: {n} operator calls done")
