"""The reference's own hot-path tests run on the CUDA path (GPU twin of
tests/test_reference_tests_on_oracle.py) plus size-independent properties at
BASELINE's full sizes."""
import hashlib
import warnings

import numpy as np
import pytest

from geotiepoints_b200 import modisinterpolator, simple_modis_interpolator, testing
from parity import haversine_m
from test_reference_tests_on_oracle import REFERENCE_LIMITS, inputs_for

pytestmark = pytest.mark.gpu

_CVIIRS_FUNCS = {
    (1000, 500): modisinterpolator.modis_1km_to_500m, (1000, 250): modisinterpolator.modis_1km_to_250m,
    (5000, 1000): modisinterpolator.modis_5km_to_1km, (5000, 500): modisinterpolator.modis_5km_to_500m,
    (5000, 250): modisinterpolator.modis_5km_to_250m,
}
_SIMPLE_FUNCS = {500: simple_modis_interpolator.modis_1km_to_500m, 250: simple_modis_interpolator.modis_1km_to_250m}


@pytest.mark.parametrize("algo,coarse,fine,width,exp,limit", REFERENCE_LIMITS)
def test_distance_limits_of_the_reference_tests(real_data, algo, coarse, fine, width, exp, limit):
    lon1, lat1, satz1 = inputs_for(real_data, coarse, width)
    with warnings.catch_warnings(record=True) as warns:
        warnings.simplefilter("always")
        if algo == "cviirs":
            lons, lats = _CVIIRS_FUNCS[(coarse, fine)](lon1, lat1, satz1)
        else:
            lons, lats = _SIMPLE_FUNCS[fine](lon1, lat1)
    assert any("may result in poor quality" in str(w.message) for w in warns) == (coarse == 5000 and fine < 1000)
    assert haversine_m(lons, lats, real_data["lon_" + exp], real_data["lat_" + exp]).max() < limit
    assert not np.isnan(lons).any() and not np.isnan(lats).any()


def test_symmetric_satz_gives_no_nan(real_data):
    satz = np.abs(np.linspace(-65.4, 65.4, 1354, dtype=np.float32)).repeat(20).reshape(-1, 20).T
    lons, lats = modisinterpolator.modis_1km_to_500m(real_data["lon_1km"], real_data["lat_1km"], np.ascontiguousarray(satz))
    assert not np.isnan(lons).any() and not np.isnan(lats).any()


def test_poles_datum(real_data):
    lon = real_data["lon_1km"] + 180
    lon = np.where(lon > 180, lon - 360, lon).astype(np.float32)
    lon5, lat5, satz5 = (np.ascontiguousarray(a[2::5, 2::5]) for a in (lon, real_data["lat_1km"], real_data["satz_1km"]))
    lons, lats = modisinterpolator.modis_5km_to_1km(lon5, lat5, satz5)
    lons = lons + 180
    lons = np.where(lons > 180, lons - 360, lons)
    assert haversine_m(lons, lats, real_data["lon_1km"], real_data["lat_1km"]).max() < 25.0


def test_zero_fixture_gives_exact_zeros():
    """BASELINE config 1: testdata/250m_lonlat_section_input.h5 holds only zeros;
    lon = lat = 0 -> xyz = (R, 0, 0) -> the output must be exactly 0 / 0."""
    lon = np.zeros((30, 1354), np.float64)
    _, _, satz = testing.modis_granule(0, n_scans=3)
    lons, lats = modisinterpolator.modis_1km_to_250m(lon, lon.copy(), satz)
    assert lons.shape == (120, 5416)
    assert np.all(lons == 0.0) and np.all(np.abs(lats) < 1e-13)
    lons, lats = simple_modis_interpolator.modis_1km_to_250m(lon, lon.copy())
    assert np.all(lons == 0.0) and np.all(np.abs(lats) < 1e-13)


# ---- size-independent properties at full size ---------------------------------------------
def _sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("dtype", [np.float32, np.float64], ids=["float32", "float64"])
def test_scans_are_independent_and_batching_is_invisible(dtype):
    """Interpolating a batch of granules in one launch, scan-range by scan-range, or with the
    scans permuted gives bit-identical pixels (what makes scan sharding across GPUs exact)."""
    parts = [testing.modis_granule(g, n_scans=203, dtype=dtype) for g in (11, 12)]
    lon, lat, satz = (np.concatenate([p[k] for p in parts]) for k in range(3))
    whole = modisinterpolator.modis_1km_to_250m(lon, lat, satz)
    assert whole[0].shape == (2 * 8120, 5416)
    assert not np.isnan(whole[0]).any() and not np.isnan(whole[1]).any()
    # determinism
    again = modisinterpolator.modis_1km_to_250m(lon, lat, satz)
    assert _sha(*whole) == _sha(*again)
    # split at an arbitrary scan boundary
    cut = 137
    a = modisinterpolator.modis_1km_to_250m(lon[:cut * 10], lat[:cut * 10], satz[:cut * 10])
    b = modisinterpolator.modis_1km_to_250m(lon[cut * 10:], lat[cut * 10:], satz[cut * 10:])
    assert _sha(np.concatenate([a[0], b[0]]), np.concatenate([a[1], b[1]])) == _sha(*whole)
    # permute scans
    rng = np.random.default_rng(0)
    perm = rng.permutation(406)
    rows_in = (perm[:, None] * 10 + np.arange(10)[None, :]).reshape(-1)
    rows_out = (perm[:, None] * 40 + np.arange(40)[None, :]).reshape(-1)
    p = modisinterpolator.modis_1km_to_250m(lon[rows_in], lat[rows_in], satz[rows_in])
    np.testing.assert_array_equal(p[0], whole[0][rows_out])
    np.testing.assert_array_equal(p[1], whole[1][rows_out])


def test_output_stays_near_the_tie_points_full_granule():
    """Every fine pixel lies within a few km of its tie-point quad (catches indexing slips
    at full size without needing the oracle)."""
    lon, lat, satz = testing.modis_granule(200, dtype=np.float64, jitter_deg=0.0)
    lons, lats = modisinterpolator.modis_1km_to_250m(lon, lat, satz)
    near_lon = np.repeat(np.repeat(lon, 4, axis=0), 4, axis=1)
    near_lat = np.repeat(np.repeat(lat, 4, axis=0), 4, axis=1)
    d = haversine_m(lons, lats, near_lon, near_lat)
    assert d.max() < 6000.0, d.max()
    lons5, lats5 = modisinterpolator.modis_5km_to_1km(*testing.subsample_5km(lon, lat, satz, 271))
    assert haversine_m(lons5, lats5, lon, lat).max() < 60.0


def _lon_diff(a, b):
    d = np.abs(a - b)
    return np.minimum(d, 360.0 - d)


@pytest.mark.parametrize("granule", [7, 14])
def test_mirror_and_rotation_symmetries_full_granule(granule):
    """Symmetries of the algorithm, at full size and without the oracle: mirroring the tie points at the
    equator (lat -> -lat) mirrors the result; rotating them about the Earth's axis (lon -> lon + 30 deg,
    the reference's own equivariance test, tests/test_modisinterpolator.py:152-164) rotates it.  Granule 14
    sweeps over the pole, so every evaluation mode of the fast path takes part."""
    lon, lat, satz = testing.modis_granule(granule)
    base = modisinterpolator.modis_1km_to_250m(lon, lat, satz)
    base = (np.array(base[0]), np.array(base[1]))
    mirrored = modisinterpolator.modis_1km_to_250m(lon, -lat, satz)
    # sin(-x) = -sin(x) exactly, so z -> -z exactly: the series around the mirrored anchor are the mirror images
    assert np.abs(mirrored[1] + base[1]).max() <= 1e-12
    assert _lon_diff(mirrored[0], base[0]).max() <= 1e-12
    shifted_in = ((lon + 30.0 + 180.0) % 360.0) - 180.0
    rotated = modisinterpolator.modis_1km_to_250m(shifted_in, lat, satz)
    assert np.abs(rotated[1] - base[1]).max() <= 1e-10
    # longitude: 1e-10 deg where it is well conditioned (tests/parity.py: away from the |sin lon| band of
    # either result and from the poles)
    expect = ((base[0] + 30.0 + 180.0) % 360.0) - 180.0
    ok = (np.abs(np.sin(np.deg2rad(base[0]))) > 2e-4) & (np.abs(np.sin(np.deg2rad(expect))) > 2e-4) & (np.abs(base[1]) < 89.9)
    assert _lon_diff(rotated[0], expect)[ok].max() <= 1e-10
    # the same two symmetries through the 5 km path and the ECEF output
    from geotiepoints_b200 import ecef
    l5 = testing.subsample_5km(lon, lat, satz, 271)
    b5 = modisinterpolator.modis_5km_to_1km(*l5)
    b5 = (np.array(b5[0]), np.array(b5[1]))
    m5 = modisinterpolator.modis_5km_to_1km(l5[0], -l5[1], l5[2])
    assert np.abs(m5[1] + b5[1]).max() <= 1e-12 and _lon_diff(m5[0], b5[0]).max() <= 1e-12
    x, y, z = (np.array(a) for a in ecef.modis_1km_to_250m(lon, lat, satz))
    xm, ym, zm = ecef.modis_1km_to_250m(lon, -lat, satz)
    assert np.abs(xm - x).max() <= 1e-8 and np.abs(ym - y).max() <= 1e-8 and np.abs(zm + z).max() <= 1e-8
