"""SURVEY.md 8(f) row 4, third item, on the GPU: geotiepoints_b200.interpolator / geointerpolator (csrc/gtp_grid.cu
through gtp_grid_interp_* / gtp_geogrid_interp_*) against the expected vectors of the reference's own tests, the
committed outputs of the reference's classes (tests/golden/grid_golden.npz), the numpy restatement
(oracle/grid_oracle.py) and - where oracle/_ref and scipy are there - the reference's classes themselves."""
import os
import sys

import numpy as np
import pytest

import grid_oracle
from conftest import GOLDEN, ROOT
from geotiepoints_b200 import testing
from geotiepoints_b200.geointerpolator import GeoGridInterpolator, GeoSplineInterpolator
from geotiepoints_b200.interpolator import (MultipleGridInterpolator, MultipleSplineInterpolator, SingleGridInterpolator,
                                            SingleSplineInterpolator)
from test_grid_oracle import (CASES, LATS_EXPECTED, LONS_EXPECTED, TIE_LATS, TIE_LONS, X_POINTS, Y_POINTS,
                              assert_lonlat_close, case_inputs, oracle_case)

pytestmark = pytest.mark.gpu


class Swath:
    """What the reference's tests pass as pyresample.geometry.SwathDefinition(lons, lats)."""

    def __init__(self, lons, lats):
        self.lons, self.lats = lons, lats


@pytest.mark.parametrize("klass,kwargs", [(GeoGridInterpolator, {}), (GeoSplineInterpolator, dict(kx=1, ky=1))])
@pytest.mark.parametrize("as_swath", [False, True])
def test_reference_test_vectors(klass, kwargs, as_swath):
    """test_geogrid_interpolation / _to_shape, test_geospline_interpolation / _to_shape (:190-232, :283-321)"""
    args = [Swath(TIE_LONS, TIE_LATS)] if as_swath else (TIE_LONS, TIE_LATS)
    interpolator = klass((Y_POINTS, X_POINTS), *args, **kwargs)
    for lons, lats in (interpolator.interpolate((np.arange(16), np.arange(8))), interpolator.interpolate_to_shape((16, 8))):
        assert lons.shape == lats.shape == (16, 8) and lons.dtype == np.float64
        np.testing.assert_allclose(lons[0, :], LONS_EXPECTED, rtol=5e-5)
        np.testing.assert_allclose(lats[:, 0], LATS_EXPECTED, rtol=5e-5)


def test_preserves_dtype_and_extrapolates():
    """test_geogrid_interpolation_preserves_dtype (:234-245), test_geogrid_interpolation_can_extrapolate (:267-277)"""
    interpolator = GeoGridInterpolator((Y_POINTS, X_POINTS), TIE_LONS.astype(np.float32), TIE_LATS.astype(np.float32))
    lons, lats = interpolator.interpolate_to_shape((16, 8))
    assert lons.dtype == np.float32 and lats.dtype == np.float32
    np.testing.assert_allclose(lons[0, :], LONS_EXPECTED, rtol=5e-5)
    interpolator = GeoGridInterpolator((Y_POINTS, X_POINTS), TIE_LONS, TIE_LATS, bounds_error=False, fill_value=None)
    lons, lats = interpolator.interpolate_to_shape((16, 16), method="cubic")
    assert lons.shape == (16, 16) and np.isfinite(lons).all() and np.isfinite(lats).all()
    exp = grid_oracle.geogrid_interpolate(Y_POINTS, X_POINTS, TIE_LONS, TIE_LATS, np.arange(16), np.arange(16), method_y=3, method_x=3)
    assert_lonlat_close((lons, lats), exp, "cubic extrapolation of the reference's test")


def product_case(name, to_device=False):
    spec = CASES[name]
    ty, tx, lon, lat, fy, fx = case_inputs(name)
    if to_device:
        import torch
        lon, lat = torch.from_numpy(lon).cuda(), torch.from_numpy(lat).cuda()
    klass = {"geogrid": GeoGridInterpolator, "geospline": GeoSplineInterpolator, "grid": MultipleGridInterpolator}[spec["kind"]]
    return tuple(klass((ty, tx), lon, lat, **spec["kwargs"]).interpolate((fy, fx)))


@pytest.mark.parametrize("name", sorted(CASES))
def test_matches_oracle_and_committed_reference_outputs(name):
    got = product_case(name)
    with np.load(os.path.join(GOLDEN, "grid_golden.npz")) as z:
        ref = z[name + "|0"], z[name + "|1"]
    for exp, what in ((oracle_case(name), "oracle"), (ref, "reference")):
        if CASES[name]["kind"] == "grid":
            for g, e in zip(got, exp):
                np.testing.assert_allclose(g, e, rtol=0, atol=1e-11, err_msg=f"{name} vs {what}")
        else:
            assert_lonlat_close(got, exp, f"{name} vs {what}")
    dev = product_case(name, to_device=True)
    for g, d in zip(got, dev):
        assert d.is_cuda and np.array_equal(d.cpu().numpy(), g, equal_nan=True), name


def test_descending_grids_and_single_classes():
    rng = np.random.default_rng(5)
    ty, tx = np.cumsum(rng.uniform(1, 3, 8)), np.cumsum(rng.uniform(1, 3, 6))
    v, w = rng.normal(size=(8, 6)), rng.normal(size=(8, 6))
    fy, fx = np.linspace(ty[0], ty[-1], 19), np.linspace(tx[0], tx[-1], 23)
    for method, m in (("linear", 1), ("nearest", 0), ("cubic", 3)):
        exp = grid_oracle.grid_interpolate(ty, tx, v, fy, fx, method_y=m, method_x=m)
        np.testing.assert_allclose(SingleGridInterpolator((ty, tx), v, method=method).interpolate((fy, fx)), exp, rtol=0, atol=1e-12)
        flipped = SingleGridInterpolator((ty[::-1], tx[::-1]), v[::-1, ::-1], method=method).interpolate((fy, fx))
        np.testing.assert_allclose(flipped, exp, rtol=0, atol=1e-12)
    for kx, ky in ((1, 1), (3, 1), (1, 3), (3, 3)):
        exp = [grid_oracle.grid_interpolate(ty, tx, a, fy, fx, method_y=kx, method_x=ky, oob=grid_oracle.CLAMP) for a in (v, w)]
        np.testing.assert_allclose(SingleSplineInterpolator((ty, tx), v, kx=kx, ky=ky).interpolate((fy, fx)), exp[0], rtol=0, atol=1e-12)
        got = list(MultipleSplineInterpolator((ty, tx), v, w, kx=kx, ky=ky).interpolate((fy, fx)))
        np.testing.assert_allclose(got[0], exp[0], rtol=0, atol=1e-12)
        np.testing.assert_allclose(got[1], exp[1], rtol=0, atol=1e-12)
    # integer tie points become float64, float32 ones stay float32 (scipy keeps floating values as they are)
    assert SingleGridInterpolator((ty, tx), (v * 10).astype(np.int32)).interpolate((fy, fx)).dtype == np.float64
    assert SingleGridInterpolator((ty, tx), v.astype(np.float32)).interpolate((fy, fx)).dtype == np.float32
    # a NaN tie point reaches the pixels of the cells around it only; a NaN coordinate gives a NaN row
    v_nan = v.copy(); v_nan[3, 2] = np.nan
    got = SingleGridInterpolator((ty, tx), v_nan).interpolate((fy, fx))
    exp = grid_oracle.grid_interpolate(ty, tx, v_nan, fy, fx)
    assert np.array_equal(np.isnan(got), np.isnan(exp)) and 0 < np.isnan(got).sum() < got.size // 2
    fy_nan = fy.copy(); fy_nan[4] = np.nan
    got = SingleGridInterpolator((ty, tx), v, bounds_error=False).interpolate((fy_nan, fx))
    assert np.isnan(got[4]).all() and np.isfinite(np.delete(got, 4, axis=0)).all()


def test_interpolate_slices_is_the_dask_task():
    """interpolator.py:290-300: what every chunk of the reference's dask graph computes."""
    interpolator = GeoGridInterpolator((Y_POINTS, X_POINTS), TIE_LONS, TIE_LATS)
    whole = interpolator.interpolate_to_shape((16, 8))
    block = interpolator._interpolate_slices([slice(0, 2), slice(4, 8), slice(4, 8)])
    assert block.shape == (2, 4, 4)
    assert np.array_equal(block[0], whole[0][4:8, 4:8]) and np.array_equal(block[1], whole[1][4:8, 4:8])
    single = SingleGridInterpolator((Y_POINTS, X_POINTS), TIE_LONS.astype(np.float64))
    assert np.array_equal(single.interpolate_slices((slice(4, 8), slice(0, 4))), single.interpolate((np.arange(16), np.arange(8)))[4:8, 0:4])


@pytest.mark.parametrize("granule", [0, 5, 144])
def test_modis_tie_points_through_the_generic_classes(granule):
    """A 5 km MODIS tie-point grid (406 x 271) to 1 km with the generic classes: the anchors of the back-transform
    are 5 km apart and sweep over the pole (granule 5) and the antimeridian (144)."""
    lon, lat, _ = testing.modis_granule(granule)
    lon5, lat5 = np.ascontiguousarray(lon[2::5, 2::5]), np.ascontiguousarray(lat[2::5, 2::5])
    ty, tx = np.arange(2, lon.shape[0], 5.0), np.arange(2, lon.shape[1], 5.0)
    fy, fx = np.arange(lon.shape[0], dtype=np.float64), np.arange(lon.shape[1], dtype=np.float64)
    for kwargs, args in ((dict(bounds_error=False, fill_value=None), dict(method_y=1, method_x=1)),
                         (dict(bounds_error=False, fill_value=None, method="cubic"), dict(method_y=3, method_x=3))):
        got = GeoGridInterpolator((ty, tx), lon5, lat5, **kwargs).interpolate((fy, fx))
        exp = grid_oracle.geogrid_interpolate(ty, tx, lon5, lat5, fy, fx, **args)
        assert_lonlat_close(got, exp, f"granule {granule} {kwargs}")


def test_matches_reference_itself():
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "geotiepoints", "geointerpolator.py")):
        pytest.skip("oracle/_ref/geotiepoints/geointerpolator.py not staged (python oracle/build_ref.py)")
    pytest.importorskip("scipy")
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    from geotiepoints import geointerpolator as ref
    lon, lat, _ = testing.modis_granule(3, n_scans=20)
    lon5, lat5 = np.ascontiguousarray(lon[2::5, 2::5]), np.ascontiguousarray(lat[2::5, 2::5])
    ty, tx = np.arange(2, lon.shape[0], 5.0), np.arange(2, lon.shape[1], 5.0)
    fy, fx = np.arange(2.0, ty[-1] + 1), np.arange(2.0, tx[-1] + 1)
    got = GeoGridInterpolator((ty, tx), lon5, lat5).interpolate((fy, fx))
    assert_lonlat_close(got, ref.GeoGridInterpolator((ty, tx), lon5, lat5).interpolate((fy, fx)), "GeoGridInterpolator")
    got = GeoSplineInterpolator((ty, tx), lon5, lat5).interpolate((fy, fx))
    assert_lonlat_close(got, ref.GeoSplineInterpolator((ty, tx), lon5, lat5).interpolate((fy, fx)), "GeoSplineInterpolator")
    got32 = GeoGridInterpolator((ty, tx), lon5.astype(np.float32), lat5.astype(np.float32)).interpolate((fy, fx))
    ref32 = ref.GeoGridInterpolator((ty, tx), lon5.astype(np.float32), lat5.astype(np.float32)).interpolate((fy, fx))
    # float32: the reference rounds the Cartesian tie points to float32 and back-transforms in float32 arithmetic - its
    # acos(x / rho) loses half the digits near lon = 0 / 180 (up to 0.02 deg here); the GPU result is the float64 one
    # rounded once.  Compared where the reference's float32 formula is well conditioned.
    ref64 = ref.GeoGridInterpolator((ty, tx), lon5, lat5).interpolate((fy, fx))
    assert got32[0].dtype == np.float32 and got32[1].dtype == np.float32
    assert np.abs(got32[0] - ref64[0]).max() <= 4e-6 and np.abs(got32[1] - ref64[1]).max() <= 4e-6      # half a float32 ulp
    well = np.abs(np.sin(np.deg2rad(ref64[0]))) > 0.5
    assert well.sum() > 1000 and np.abs(got32[0] - ref32[0])[well].max() <= 3e-5 and np.abs(got32[1] - ref32[1]).max() <= 3e-5


def _fake_dask_modules():
    """dask is not in the image: just enough of dask.base / dask.array / dask.array.core for the graphs that
    interpolator.py:261-284 builds ({(name, i, j): (func, slices)} handed to da.Array)."""
    import types

    def normalize_chunks(chunks, shape, dtype=None):
        if chunks == "auto":
            chunks = shape
        if isinstance(chunks, int):
            chunks = (chunks,) * len(shape)
        out = []
        for c, n in zip(chunks, shape):
            if isinstance(c, (tuple, list)):
                out.append(tuple(c))
            else:
                full, rest = divmod(n, c)
                out.append((c,) * full + ((rest,) if rest else ()))
        return tuple(out)

    class Array:
        def __init__(self, dsk, name, shape, chunks, dtype, lead=None):
            self.dask, self.name, self.shape, self.chunks, self.dtype, self._lead = dsk, name, tuple(shape), tuple(chunks), np.dtype(dtype), lead

        def __getitem__(self, k):
            assert isinstance(k, int) and self._lead is None and self.chunks[0] == (self.shape[0],)
            return Array(self.dask, self.name, self.shape[1:], self.chunks[1:], self.dtype, lead=k)

        def compute(self):
            chunks = self.chunks if self._lead is None else ((1,),) + self.chunks
            rows = []
            for i in range(len(chunks[-2])):
                row = []
                for j in range(len(chunks[-1])):
                    key = (self.name,) + ((0,) if self._lead is not None else ()) + (i, j)
                    func, arg = self.dask[key]
                    block = np.asarray(func(arg))
                    row.append(block if self._lead is None else block[self._lead])
                rows.append(np.concatenate(row, axis=-1))
            return np.concatenate(rows, axis=-2)

    base = types.ModuleType("dask.base"); base.tokenize = lambda *a, **k: format(abs(hash(repr((a, k)))), "x")
    core = types.ModuleType("dask.array.core"); core.normalize_chunks = normalize_chunks
    array = types.ModuleType("dask.array"); array.Array = Array; array.core = core
    dask = types.ModuleType("dask"); dask.base = base; dask.array = array
    return {"dask": dask, "dask.base": base, "dask.array": array, "dask.array.core": core}


def test_chunked_interpolation_builds_the_references_graph(monkeypatch):
    """test_chunked_geogrid_interpolation / test_chunked_geospline_interpolation (:247-265, :338-356)"""
    for name, mod in _fake_dask_modules().items():
        monkeypatch.setitem(sys.modules, name, mod)
    lons32, lats32 = TIE_LONS.astype(np.float32), TIE_LATS.astype(np.float32)
    interpolator = GeoGridInterpolator((Y_POINTS, X_POINTS), lons32, lats32)
    eager = interpolator.interpolate_to_shape((16, 8))
    lons, lats = interpolator.interpolate_to_shape((16, 8), chunks=4)
    assert lons.chunks == ((4, 4, 4, 4), (4, 4)) and lats.chunks == ((4, 4, 4, 4), (4, 4)) and lons.dtype == np.float32
    assert np.array_equal(lons.compute(), eager[0]) and np.array_equal(lats.compute(), eager[1])
    lons, lats = interpolator.interpolate_to_shape((16, 8), chunks=(8, 3))
    assert lons.chunks == ((8, 8), (3, 3, 2)) and np.array_equal(lats.compute(), eager[1])
    multi = MultipleGridInterpolator((Y_POINTS, X_POINTS), lons32, lats32)
    a, b = multi.interpolate_to_shape((16, 8), chunks=4)
    ea, eb = multi.interpolate_to_shape((16, 8))
    assert a.chunks == ((4, 4, 4, 4), (4, 4)) and np.array_equal(a.compute(), ea) and np.array_equal(b.compute(), eb)


def test_concurrent_calls_from_threads():
    """dask's threaded scheduler runs the chunk tasks of one interpolator concurrently (threading)."""
    from concurrent.futures import ThreadPoolExecutor
    lon, lat, _ = testing.modis_granule(7, n_scans=10)
    lon5, lat5 = np.ascontiguousarray(lon[2::5, 2::5]), np.ascontiguousarray(lat[2::5, 2::5])
    ty, tx = np.arange(2, lon.shape[0], 5.0), np.arange(2, lon.shape[1], 5.0)
    interpolator = GeoGridInterpolator((ty, tx), lon5, lat5, bounds_error=False, fill_value=None)
    whole = interpolator.interpolate_to_shape(lon.shape)
    pieces = [(slice(0, 2), slice(r, r + 10), slice(c, c + 677)) for r in range(0, 100, 10) for c in (0, 677)]
    with ThreadPoolExecutor(8) as pool:
        blocks = list(pool.map(interpolator._interpolate_slices, pieces))
    for (_, rows, cols), block in zip(pieces, blocks):
        assert np.array_equal(block[0], whole[0][rows, cols]) and np.array_equal(block[1], whole[1][rows, cols])
