"""The reference's own hot-path tests, re-expressed against the oracle (CPU).

geotiepoints/tests/test_modisinterpolator.py:110-164 and
test_simple_modis_interpolator.py:21-53 pin the path with geodesic-distance limits against
testdata/modis_test_data.h5.  h5py / pyproj / dask / xarray are not installed, so the file
is read from the re-packed npz and distances are great-circle (tests/parity.py).
The same table is run on the GPU in tests/test_gpu_reference_tests.py.
"""
import numpy as np
import pytest

import gtp_oracle
from parity import haversine_m

# (algorithm, coarse, fine, width, expected dataset suffix, limit in metres) - the limits of
# tests/test_modisinterpolator.py:113-120 and tests/test_simple_modis_interpolator.py:24-29
REFERENCE_LIMITS = [
    ("cviirs", 1000, 500, 0, "500m", 5.0),
    ("cviirs", 1000, 250, 0, "250m", 8.30),
    ("cviirs", 5000, 1000, 271, "1km", 25.0),
    ("cviirs", 5000, 1000, 270, "1km", 110.0),
    ("cviirs", 5000, 500, 271, "500m", 19500.0),
    ("cviirs", 5000, 250, 271, "250m", 25800.0),
    ("simple", 1000, 500, 0, "500m", 16.0),
    ("simple", 1000, 250, 0, "250m", 27.35),
]


def inputs_for(real_data, coarse, width):
    lon, lat, satz = real_data["lon_1km"], real_data["lat_1km"], real_data["satz_1km"]
    if coarse == 5000:
        sl = (slice(2, None, 5), slice(2, None, 5)) if width == 271 else (slice(2, None, 5), slice(2, -5, 5))
        return tuple(np.ascontiguousarray(a[sl]) for a in (lon, lat, satz))
    return lon, lat, satz


@pytest.mark.parametrize("algo,coarse,fine,width,exp,limit", REFERENCE_LIMITS)
def test_distance_limits(real_data, algo, coarse, fine, width, exp, limit):
    lon1, lat1, satz1 = inputs_for(real_data, coarse, width)
    if algo == "cviirs":
        lons, lats = gtp_oracle.cviirs(lon1, lat1, satz1, coarse, fine, width)
    else:
        lons, lats = gtp_oracle.simple(lon1, lat1, coarse, fine)
    dist = haversine_m(lons, lats, real_data["lon_" + exp], real_data["lat_" + exp])
    assert dist.max() < limit
    assert not np.isnan(lons).any() and not np.isnan(lats).any()


def test_symmetric_satz_gives_no_nan(real_data):
    """GH #19, tests/test_modisinterpolator.py:143-149."""
    satz = np.abs(np.linspace(-65.4, 65.4, 1354, dtype=np.float32)).repeat(20).reshape(-1, 20).T
    lons, lats = gtp_oracle.cviirs(real_data["lon_1km"], real_data["lat_1km"], np.ascontiguousarray(satz), 1000, 500)
    assert not np.isnan(lons).any() and not np.isnan(lats).any()


def test_poles_datum(real_data):
    """tests/test_modisinterpolator.py:152-164: shifting lon by 180 deg commutes with the interpolation."""
    lon = real_data["lon_1km"] + 180
    lon = np.where(lon > 180, lon - 360, lon).astype(np.float32)
    lon5, lat5, satz5 = (np.ascontiguousarray(a[2::5, 2::5]) for a in (lon, real_data["lat_1km"], real_data["satz_1km"]))
    lons, lats = gtp_oracle.cviirs(lon5, lat5, satz5, 5000, 1000, 271)
    lons = lons + 180
    lons = np.where(lons > 180, lons - 360, lons)
    assert haversine_m(lons, lats, real_data["lon_1km"], real_data["lat_1km"]).max() < 25.0
