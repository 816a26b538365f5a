#!/usr/bin/env python
"""Golden fixtures of the LEGACY spline interpolators, generated from the REAL reference.

Run in the build container only (needs /root/reference and oracle/_ref built by ``python oracle/build_ref.py``):

    python tests/golden/make_legacy_golden.py

Writes ``legacy_golden.npz`` beside this script:
  ref5km/lon, ref5km/lat, full1km/lon, full1km/lat
      the reference's own golden pair testdata/test_5_to_1_geoloc_{5km,full}.h5 (int32 millidegrees,
      (10, 271) -> (50, 1354); geotiepoints/tests/test_modis.py:41-55), read with tests/golden/h5lite.py
  <kind>|g<granule>|lon, ...|lat
      a sample (every row; every 53rd column plus the first / last 6) of what the reference's
      modis5kmto1km / modis1kmto500m / modis1kmto250m (geotiepoints/__init__.py:49-152, scipy
      RectBivariateSpline) return for ``testing.modis_granule(granule, n_scans=2)`` (5 km: subsampled
      [2::5, 2::5]) - the inputs are deterministic, so only the outputs are stored.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, os.path.join(ROOT, "python-geotiepoints_b200"))

from h5lite import H5File  # noqa: E402
from geotiepoints_b200 import testing  # noqa: E402
from geotiepoints import legacy  # noqa: E402  (the reference's geotiepoints/__init__.py, see oracle/build_ref.py)

from make_legacy_golden_inputs import GRANULES, legacy_inputs, sample_cols  # noqa: E402


def main():
    out = {}
    for key, fname in (("ref5km", "test_5_to_1_geoloc_5km.h5"), ("full1km", "test_5_to_1_geoloc_full.h5")):
        h5 = H5File(os.path.join("/root/reference/testdata", fname))
        out[key + "/lon"], out[key + "/lat"] = h5["longitude"], h5["latitude"]
    fns = {"5km_1km": legacy.modis5kmto1km, "1km_500m": legacy.modis1kmto500m, "1km_250m": legacy.modis1kmto250m}
    for kind, fn in fns.items():
        for g in GRANULES:
            with np.errstate(all="ignore"):
                lons, lats = fn(*legacy_inputs(g, kind))
            cols = sample_cols(lons.shape[1])
            out[f"{kind}|g{g}|lon"], out[f"{kind}|g{g}|lat"] = lons[:, cols], lats[:, cols]
    np.savez_compressed(os.path.join(HERE, "legacy_golden.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
