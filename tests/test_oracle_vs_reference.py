"""The oracle against the reference itself (oracle/_ref, built by oracle/build_ref.py).

Runs wherever the built reference is importable (the build container, and the GPU box,
where oracle/_ref travels with the snapshot); skipped otherwise.  float32 and the simple
path use the UNMODIFIED reference; float64 CVIIRS uses the twin with the 2-line dtype
patch the stock code needs to accept float64 at all.
"""
import os
import sys

import numpy as np
import pytest

import gtp_oracle
from cases import CVIIRS_CONFIGS, SIMPLE_CONFIGS
from conftest import ROOT
from geotiepoints_b200 import testing

REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def _import_ref():
    if not os.path.isdir(os.path.join(REF_DIR, "geotiepoints")):
        pytest.skip("oracle/_ref not built (python oracle/build_ref.py)")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    try:
        import geotiepoints._modis_interpolator as stock
        import geotiepoints_f64._modis_interpolator as f64
        import geotiepoints._simple_modis_interpolator as simple
    except ImportError as exc:  # e.g. a different Python ABI
        pytest.skip(f"oracle/_ref not importable here: {exc}")
    return stock, f64, simple


@pytest.mark.parametrize("granule", [3, 5, 41])
def test_cviirs_bit_exact(granule):
    stock, f64, _ = _import_ref()
    lon, lat, satz = testing.modis_granule(granule, n_scans=3)
    for dtype, ref in ((np.float32, stock), (np.float64, f64)):
        a1 = tuple(a.astype(dtype) for a in (lon, lat, satz))
        for coarse, fine, width in CVIIRS_CONFIGS:
            a = testing.subsample_5km(*a1, width) if coarse == 5000 else a1
            r = ref.interpolate(*a, coarse_resolution=coarse, fine_resolution=fine, coarse_scan_width=width)
            o = gtp_oracle.cviirs(*a, coarse, fine, width)
            np.testing.assert_array_equal(r[0], o[0])
            np.testing.assert_array_equal(r[1], o[1])


@pytest.mark.parametrize("granule", [3, 5, 41])
def test_simple_bit_exact(granule):
    _, _, simple = _import_ref()
    lon, lat, _ = testing.modis_granule(granule, n_scans=3)
    for dtype in (np.float32, np.float64):
        for coarse, fine in SIMPLE_CONFIGS:
            a = lon.astype(dtype), lat.astype(dtype)
            r = simple.interpolate_geolocation_cartesian(*a, coarse, fine)
            o = gtp_oracle.simple(*a, coarse, fine)
            np.testing.assert_array_equal(r[0], o[0])
            np.testing.assert_array_equal(r[1], o[1])


def test_stock_reference_rejects_float64():
    """the stock CVIIRS path raises for float64 input."""
    stock, _, _ = _import_ref()
    lon, lat, satz = testing.modis_granule(0, n_scans=1)
    with pytest.raises(ValueError, match="Buffer dtype mismatch"):
        stock.interpolate(lon, lat, satz, coarse_resolution=1000, fine_resolution=250)


def test_simple_nonstandard_width():
    _, _, simple = _import_ref()
    lon, lat, _ = testing.modis_granule(7, n_scans=2)
    a = np.ascontiguousarray(lon[:, :100]), np.ascontiguousarray(lat[:, :100])
    r = simple.interpolate_geolocation_cartesian(*a, 1000, 250)
    o = gtp_oracle.simple(*a, 1000, 250)
    np.testing.assert_array_equal(r[0], o[0])
    np.testing.assert_array_equal(r[1], o[1])
