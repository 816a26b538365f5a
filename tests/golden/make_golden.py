#!/usr/bin/env python
"""Generate the committed golden fixtures from the REAL reference.

Run in the build container only (needs /root/reference and oracle/_ref built by
``python oracle/build_ref.py``):

    python tests/golden/make_golden.py

Writes, beside this script:
  modis_test_data.npz   the reference's own golden file testdata/modis_test_data.h5
                        (geotiepoints/tests/test_modisinterpolator.py:18-19) re-packed as
                        npz (h5py is not installed; read with tests/golden/h5lite.py)
  stress_inputs.npz     one-scan float32 stress granules (pole, antimeridian, ...)
  reference_outputs.json  sha256 of the reference's lon/lat output bytes (NaNs
                        canonicalised) for every (input, algorithm, resolution, width,
                        dtype) case, produced by the UNMODIFIED reference for float32 and
                        for the simple path, and by the 2-line-patched f64 twin for the
                        float64 CVIIRS cases - `impl` says which
  reference_samples.npz a strided sample of each output (rows ::5, columns ::97 plus
                        the 8 first / last columns) for readable failure messages
"""
import hashlib
import json
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

import geotiepoints._modis_interpolator as ref_cviirs_stock  # noqa: E402
import geotiepoints_f64._modis_interpolator as ref_cviirs_f64  # noqa: E402
import geotiepoints._simple_modis_interpolator as ref_simple  # noqa: E402

REFERENCE_TESTDATA = "/root/reference/testdata"

CVIIRS_CONFIGS = [  # (coarse, fine, width)
    (1000, 250, 0), (1000, 500, 0),
    (5000, 1000, 271), (5000, 500, 271), (5000, 250, 271),
    (5000, 1000, 270), (5000, 500, 270), (5000, 250, 270),
]
SIMPLE_CONFIGS = [(1000, 250), (1000, 500)]


def canonical_bytes(a):
    a = np.ascontiguousarray(a).copy()
    a[np.isnan(a)] = np.nan          # one canonical quiet-NaN bit pattern
    return a.tobytes()


def digest(a):
    return hashlib.sha256(canonical_bytes(a)).hexdigest()


def sample(a):
    cols = np.unique(np.concatenate([np.arange(0, a.shape[1], 97), np.arange(8), np.arange(a.shape[1] - 8, a.shape[1])]))
    return a[::5][:, cols]


def case_id(inp, algo, coarse, fine, width, dtype):
    return f"{inp}|{algo}|{coarse}->{fine}|w{width}|{np.dtype(dtype).name}"


def all_inputs():
    """name -> (lon, lat, satz) float32 1 km arrays."""
    h5 = H5File(os.path.join(REFERENCE_TESTDATA, "modis_test_data.h5"))
    real = {k: h5[k] for k in h5.keys()}
    np.savez_compressed(os.path.join(HERE, "modis_test_data.npz"), **real)
    inputs = {"real": (real["lon_1km"], real["lat_1km"], real["satz_1km"])}
    stress = {}
    for kind in testing.STRESS_KINDS:
        lon, lat, satz = testing.stress_granule(kind, n_scans=1, dtype=np.float32)
        stress[kind] = (lon, lat, satz)
    stress["nan_holes"] = testing.punch_nan_holes(*inputs["real"])
    flat = {}
    for kind, (lon, lat, satz) in stress.items():
        flat[kind + "/lon"], flat[kind + "/lat"], flat[kind + "/satz"] = lon, lat, satz
    np.savez_compressed(os.path.join(HERE, "stress_inputs.npz"), **flat)
    inputs.update(stress)
    return inputs


def main():
    inputs = all_inputs()
    digests, samples = {}, {}
    for name, arrs32 in inputs.items():
        for dtype in (np.float32, np.float64):
            lon, lat, satz = (a.astype(dtype) for a in arrs32)
            ref = ref_cviirs_stock if dtype == np.float32 else ref_cviirs_f64
            impl = "reference (stock)" if dtype == np.float32 else "reference + 2-line f64 dtype patch"
            for coarse, fine, width in CVIIRS_CONFIGS:
                if coarse == 5000:
                    a = testing.subsample_5km(lon, lat, satz, width)
                else:
                    a = (lon, lat, satz)
                with np.errstate(all="ignore"):
                    olon, olat = ref.interpolate(*a, coarse_resolution=coarse, fine_resolution=fine,
                                                 coarse_scan_width=width)
                cid = case_id(name, "cviirs", coarse, fine, width, dtype)
                digests[cid] = {"lon_sha256": digest(olon), "lat_sha256": digest(olat),
                                "shape": list(olon.shape), "impl": impl,
                                "nan_count": int(np.isnan(olon).sum() + np.isnan(olat).sum())}
                samples[cid + "|lon"], samples[cid + "|lat"] = sample(olon), sample(olat)
            for coarse, fine in SIMPLE_CONFIGS:
                with np.errstate(all="ignore"):
                    olon, olat = ref_simple.interpolate_geolocation_cartesian(lon, lat, coarse, fine)
                cid = case_id(name, "simple", coarse, fine, 0, dtype)
                digests[cid] = {"lon_sha256": digest(olon), "lat_sha256": digest(olat),
                                "shape": list(olon.shape), "impl": "reference (stock) + scipy " + _scipy_version(),
                                "nan_count": int(np.isnan(olon).sum() + np.isnan(olat).sum())}
                samples[cid + "|lon"], samples[cid + "|lat"] = sample(olon), sample(olat)
    with open(os.path.join(HERE, "reference_outputs.json"), "w") as fh:
        json.dump(digests, fh, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "reference_samples.npz"), **samples)
    print(f"{len(digests)} cases written")


def _scipy_version():
    import scipy
    return scipy.__version__


if __name__ == "__main__":
    main()
