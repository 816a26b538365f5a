"""The numpy restatement of the generic grid interpolators (oracle/grid_oracle.py) against the reference
(third item), and the host logic of geotiepoints_b200.interpolator / geointerpolator.

Pins: the expected vectors of the reference's own tests (geotiepoints/tests/test_geointerpolator.py:187-278);
outputs of the reference's classes (scipy RegularGridInterpolator / RectBivariateSpline underneath) committed in
tests/golden/grid_golden.npz by make_grid_golden.py; the reference's classes themselves where oracle/_ref and
scipy are present.
"""
import os
import sys

import numpy as np
import pytest

import grid_oracle
from conftest import GOLDEN, ROOT

sys.path.insert(0, GOLDEN)
from grid_cases import CASES, case_inputs  # noqa: E402

METHODS = {"linear": 1, "slinear": 1, "nearest": 0, "cubic": 3}

# geotiepoints/tests/test_geointerpolator.py:174-184, :201-205
TIE_LONS = np.array([[1, 2, 3, 4]] * 5)
TIE_LATS = np.array([[1] * 4, [2] * 4, [3] * 4, [4] * 4, [5] * 4])
X_POINTS, Y_POINTS = np.array([0, 1, 3, 7]), np.array([0, 1, 3, 7, 15])
LONS_EXPECTED = np.array([1., 2., 2.5, 3., 3.25, 3.5, 3.75, 4.])
LATS_EXPECTED = np.array([1., 2., 2.5, 3., 3.25, 3.5, 3.75, 4., 4.125, 4.25, 4.375, 4.5, 4.625, 4.75, 4.875, 5.])


def oracle_case(name):
    """What the oracle computes for a case of grid_cases.py: two arrays."""
    spec = CASES[name]
    ty, tx, lon, lat, fy, fx = case_inputs(name)
    kw = spec["kwargs"]
    if spec["kind"] == "geospline":
        args = dict(method_y=kw.get("kx", 3), method_x=kw.get("ky", 3), oob=grid_oracle.CLAMP)
    else:
        m = METHODS[kw.get("method", "linear")]
        extrapolate = kw.get("bounds_error", True) or kw.get("fill_value", np.nan) is None
        args = dict(method_y=m, method_x=m, oob=grid_oracle.EXTRAPOLATE if extrapolate else grid_oracle.FILL,
                    fill=np.nan if extrapolate else kw.get("fill_value", np.nan))
    if spec["kind"] == "grid":
        return tuple(grid_oracle.grid_interpolate(ty, tx, v, fy, fx, **args) for v in (lon, lat))
    return grid_oracle.geogrid_interpolate(ty, tx, lon, lat, fy, fx, **args)


def assert_lonlat_close(got, exp, name, tol=1e-10):
    """lon / lat within tol degrees; longitudes compared on the circle and not where lat is within 1e-3 deg of a pole
    or the longitude's cosine is within 2e-4 of +-1 (acos conditioning of the reference's own formula)."""
    for k in (0, 1):
        assert got[k].shape == exp[k].shape, name
        assert np.array_equal(np.isnan(got[k]), np.isnan(exp[k])), name
    dlat = np.abs(got[1] - exp[1])
    dlon = np.abs(got[0] - exp[0])
    dlon = np.minimum(dlon, 360.0 - dlon)
    ill = (np.abs(np.sin(np.deg2rad(exp[0]))) < 2e-4) | (np.abs(exp[1]) > 89.999)
    assert np.nanmax(dlat, initial=0.0) <= tol, (name, np.nanmax(dlat))
    assert np.nanmax(np.where(ill, 0.0, dlon), initial=0.0) <= tol, (name, np.nanmax(np.where(ill, 0.0, dlon)))
    assert np.nanmax(dlon, initial=0.0) <= 2e-6, (name, np.nanmax(dlon))


def test_reference_test_vectors_oracle():
    """test_geogrid_interpolation / test_geospline_interpolation (:190-209, :283-302): rtol 5e-5."""
    for args in (dict(method_y=1, method_x=1), dict(method_y=1, method_x=1, oob=grid_oracle.CLAMP)):
        lons, lats = grid_oracle.geogrid_interpolate(Y_POINTS, X_POINTS, TIE_LONS, TIE_LATS, np.arange(16), np.arange(8), **args)
        np.testing.assert_allclose(lons[0, :], LONS_EXPECTED, rtol=5e-5)
        np.testing.assert_allclose(lats[:, 0], LATS_EXPECTED, rtol=5e-5)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_committed_reference_outputs(name):
    with np.load(os.path.join(GOLDEN, "grid_golden.npz")) as z:
        exp = z[name + "|0"], z[name + "|1"]
    got = oracle_case(name)
    if CASES[name]["kind"] == "grid":
        for g, e in zip(got, exp):
            np.testing.assert_allclose(g, e, rtol=0, atol=1e-12)
    else:
        assert_lonlat_close(got, exp, name)


def test_oracle_matches_reference_itself():
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "geotiepoints", "geointerpolator.py")):
        pytest.skip("oracle/_ref/geotiepoints/geointerpolator.py not staged (python oracle/build_ref.py)")
    pytest.importorskip("scipy")
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    from geotiepoints.geointerpolator import GeoGridInterpolator
    rng = np.random.default_rng(11)
    ty, tx = np.cumsum(rng.uniform(1, 4, 15)), np.cumsum(rng.uniform(1, 4, 13))
    lon = -20 + 0.02 * tx[None, :] + 0.004 * ty[:, None]
    lat = 30 + 0.015 * ty[:, None] - 0.003 * tx[None, :]
    fy, fx = np.linspace(ty[0], ty[-1], 57), np.linspace(tx[0], tx[-1], 44)
    ref = GeoGridInterpolator((ty, tx), lon, lat)
    assert_lonlat_close(grid_oracle.geogrid_interpolate(ty, tx, lon, lat, fy, fx), ref.interpolate((fy, fx)), "linear")
    # the reference's default "cubic" carries the tolerance of scipy's iterative spline construction (gcrotmk, 1e-5
    # relative): ~1e-4 deg here; with a direct solver it is the restatement's spline to rounding
    import scipy.sparse.linalg as ssl
    loose = GeoGridInterpolator((ty, tx), lon, lat, method="cubic").interpolate((fy, fx))
    tight = GeoGridInterpolator((ty, tx), lon, lat, method="cubic", solver=ssl.spsolve).interpolate((fy, fx))
    got = grid_oracle.geogrid_interpolate(ty, tx, lon, lat, fy, fx, method_y=3, method_x=3)
    assert_lonlat_close(got, tight, "cubic, direct solver")
    assert np.abs(got[0] - loose[0]).max() < 5e-3 and np.abs(got[1] - loose[1]).max() < 5e-3


# ---- host logic of the product classes (no GPU: everything here fails or returns before the library is needed) ----

def test_geogrid_counts_its_arguments():
    """test_geogrid_interpolation_counts_its_arguments (:211-214)"""
    from geotiepoints_b200.geointerpolator import GeoGridInterpolator
    with pytest.raises(ValueError, match="Either pass lon/lats or a pyresample definition."):
        GeoGridInterpolator((None, None), None, None, None)


def test_grid_validation_like_scipy():
    from geotiepoints_b200.geointerpolator import GeoGridInterpolator, GeoSplineInterpolator
    from geotiepoints_b200.interpolator import MultipleGridInterpolator, SingleGridInterpolator
    with pytest.raises(ValueError, match="Method 'spam' is not defined"):
        GeoGridInterpolator((Y_POINTS, X_POINTS), TIE_LONS, TIE_LATS, method="spam")
    with pytest.raises(ValueError, match="strictly ascending or descending"):
        GeoGridInterpolator((np.array([0, 1, 1, 7, 15]), X_POINTS), TIE_LONS, TIE_LATS)
    with pytest.raises(ValueError, match="There are 3 points and 4 values in dimension 1"):
        SingleGridInterpolator((Y_POINTS, X_POINTS[:3]), TIE_LONS)
    itp = GeoGridInterpolator((Y_POINTS, X_POINTS), TIE_LONS, TIE_LATS)
    with pytest.raises(ValueError, match="One of the requested xi is out of bounds in dimension 1"):
        itp.interpolate_to_shape((16, 16))
    with pytest.raises(ValueError, match="Method 'spam' is not defined"):
        itp.interpolate_to_shape((16, 8), method="spam")
    with pytest.raises(NotImplementedError):
        itp.interpolate_to_shape((16, 8), method="pchip")
    with pytest.raises(NotImplementedError):
        GeoSplineInterpolator((Y_POINTS, X_POINTS), TIE_LONS, TIE_LATS, kx=2)
    with pytest.raises(NotImplementedError):
        GeoSplineInterpolator((Y_POINTS, X_POINTS), TIE_LONS, TIE_LATS, s=1.0)
    with pytest.raises(ValueError, match="requires at least"):
        MultipleGridInterpolator((np.arange(3), X_POINTS), TIE_LONS[:3], bounds_error=False).interpolate(
            (np.arange(3), np.arange(4)), method="cubic").__next__()


def test_chunk_slices_enumeration():
    """interpolator.py:303-312"""
    from geotiepoints_b200.interpolator import _enumerate_chunk_slices
    got = list(_enumerate_chunk_slices(((4, 4), (3, 5))))
    assert got == [((0, 0), [slice(0, 4), slice(0, 3)]), ((0, 1), [slice(0, 4), slice(3, 8)]),
                   ((1, 0), [slice(4, 8), slice(0, 3)]), ((1, 1), [slice(4, 8), slice(3, 8)])]
