"""SURVEY.md 8(f) row 4, second item, on the GPU: geotiepoints_b200.viiinterpolator (csrc/gtp_vii.cu) against the
golden vectors of the reference's own tests (tests/vii_golden.py) and against the numpy restatement
(oracle/vii_oracle.py) on METimage-sized granules."""
import numpy as np
import pytest

import vii_golden
import vii_oracle
from geotiepoints_b200 import testing, viiinterpolator

pytestmark = pytest.mark.gpu


def test_reference_test_vectors():
    """geotiepoints/tests/test_viiinterpolator.py:160-222 through the CUDA path (numpy in / out)."""
    vii_golden.check_reference_cases(viiinterpolator.tie_points_interpolation, viiinterpolator.tie_points_geo_interpolation)


def _vii_like_tie_points(granule, n_scans=6, scan_alt=4, act=421):
    """Tie points with the structure of a VII granule (every scan has its own edge tie rows): MODIS-like orbit
    geometry sub-sampled to `scan_alt` rows per scan and `act` columns."""
    lon, lat, _ = testing.modis_granule(granule, n_scans=n_scans)
    rows = np.concatenate([s * 10 + np.linspace(0, 9, scan_alt).round().astype(int) for s in range(n_scans)])
    cols = np.linspace(0, 1353, act).round().astype(int)
    return np.ascontiguousarray(lon[np.ix_(rows, cols)]), np.ascontiguousarray(lat[np.ix_(rows, cols)])


@pytest.mark.parametrize("granule", [0, 3, 5, 144])          # geodetic, geodetic, polar (Cartesian), antimeridian (Cartesian)
def test_vii_geo_matches_oracle(granule):
    import torch
    lon, lat = _vii_like_tie_points(granule)
    exp = vii_oracle.tie_points_geo_interpolation(lon, lat, 4, 8)
    got = viiinterpolator.tie_points_geo_interpolation(lon, lat, 4, 8)
    assert got[0].shape == exp[0].shape == (6 * 3 * 8, 420 * 8) and got[0].dtype == np.float64
    cart = vii_oracle.use_cartesian(lon, lat)
    tol = 1e-10 if not cart else 1e-9          # Cartesian: library acos next to lon = +-180 / the poles (conditioning)
    dlon = np.abs(got[0] - exp[0]); dlon = np.minimum(dlon, 360.0 - dlon)
    band = np.abs(np.sin(np.deg2rad(exp[0]))) < 2e-4
    assert np.nanmax(np.abs(got[1] - exp[1])) <= 1e-10 and np.nanmax(np.where(band, 0.0, dlon)) <= tol, (granule, cart)
    dev = viiinterpolator.tie_points_geo_interpolation(torch.from_numpy(lon).cuda(), torch.from_numpy(lat).cuda(), 4, 8)
    assert dev[0].is_cuda and np.array_equal(dev[0].cpu().numpy(), got[0], equal_nan=True) and np.array_equal(dev[1].cpu().numpy(), got[1], equal_nan=True)


@pytest.mark.parametrize("factor", [3, 4])          # 3: the per-pixel kernel; 4 (and 8 above): a thread per tie cell
def test_vii_geo_other_factors(factor):
    lon, lat = _vii_like_tie_points(5)             # polar: Cartesian mode
    exp = vii_oracle.tie_points_geo_interpolation(lon, lat, 4, factor)
    got = viiinterpolator.tie_points_geo_interpolation(lon, lat, 4, factor)
    assert got[0].shape == exp[0].shape == (6 * 3 * factor, 420 * factor)
    dlon = np.abs(got[0] - exp[0]); dlon = np.minimum(dlon, 360.0 - dlon)
    band = np.abs(np.sin(np.deg2rad(exp[0]))) < 2e-4
    assert np.nanmax(np.abs(got[1] - exp[1])) <= 1e-10 and np.nanmax(np.where(band, 0.0, dlon)) <= 1e-9


def test_vii_generic_arrays_and_dataarray_like():
    rng = np.random.default_rng(3)
    a, b = rng.normal(size=(12, 33)), rng.normal(size=(12, 33))
    exp = vii_oracle.tie_points_interpolation([a, b], 4, 8)
    got = viiinterpolator.tie_points_interpolation([a, b], 4, 8)
    assert all(np.allclose(g, e, rtol=0, atol=1e-14) for g, e in zip(got, exp))

    class DataArrayLike:
        def __init__(self, values, dims):
            self.values, self.dims, self.shape = np.asarray(values), tuple(dims), np.asarray(values).shape

    res = viiinterpolator.tie_points_interpolation([DataArrayLike(a, ("alt", "act"))], 4, 8)[0]
    assert isinstance(res, DataArrayLike) and res.dims == ("alt", "act") and np.allclose(res.values, exp[0], atol=1e-14)
    with pytest.raises(ValueError, match="not consistent"):
        viiinterpolator.tie_points_interpolation([a, b[:, :-1]], 4, 8)
