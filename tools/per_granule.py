#!/usr/bin/env python
"""Kernel time of the CVIIRS 1 km -> 250 m float64 path per synthetic granule (which latitudes cost
what):  GTP_B200_LIB=<lib.so> python tools/per_granule.py 0 3 4 14 ..."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "python-geotiepoints_b200"))
import numpy as np
import torch
from geotiepoints_b200 import _capi, testing

lib = _capi.lib()
dev = torch.device("cuda", 0)
reps = 4          # granule repeated to make the launch long enough to time
for g in [int(a) for a in sys.argv[1:]] or [0, 4]:
    arrs = [np.concatenate([a] * reps) for a in testing.modis_granule(g)]
    d_in = [torch.from_numpy(a).to(dev) for a in arrs]
    n_scans = arrs[0].shape[0] // 10
    out = [torch.empty((n_scans * 40, 5416), dtype=torch.float64, device=dev) for _ in range(2)]
    st = torch.cuda.current_stream(dev)

    def step():
        _capi.check(lib.gtp_cviirs_interp_f64(*[t.data_ptr() for t in d_in], n_scans, 1000, 250, 0,
                                              out[0].data_ptr(), out[1].data_ptr(), 0, ctypes.c_void_p(st.cuda_stream)))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5):
        step()
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    px = n_scans * 40 * 5416
    print(f"granule {g:3d} lat {arrs[1].min():6.1f}..{arrs[1].max():5.1f}  {ms:7.3f} ms  {px / ms / 1e6:8.1f} Gpx/s  "
          f"{17.5 * px / ms / 1e6 / 6547.8:5.3f} of roofline")
