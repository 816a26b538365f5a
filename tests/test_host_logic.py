"""Host-side logic that needs no GPU: the array-container adapter, argument validation
with the reference's error types and messages, the C-ABI export list, scan sharding."""
import ctypes
import os
import re
import types
import warnings

import numpy as np
import pytest

import fake_dask
from conftest import ROOT
from geotiepoints_b200 import _capi, _modis_utils, modisinterpolator, sharding, simple_modis_interpolator
from geotiepoints_b200._modis_interpolator import MODISInterpolator, interpolate
from geotiepoints_b200._modis_utils import rows_per_scan_for_resolution, scanline_mapblocks


# ---- C ABI ---------------------------------------------------------------------------
def _header_functions():
    text = open(os.path.join(ROOT, "include", "gtp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gtp_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = _header_functions()
    assert len(names) >= 16
    handle = ctypes.CDLL(_capi.LIB_PATH)
    for name in names:
        assert hasattr(handle, name), f"{name} declared in include/gtp_b200.h but not exported"
    assert set(names) == set(_capi.SIGNATURES), "ctypes binding and header disagree"


def test_version_and_geometry_without_gpu():
    lib = _capi.lib()
    assert b"gtp_b200" in lib.gtp_version()
    vals = [ctypes.c_int() for _ in range(4)]
    for (coarse, fine, width), want in {
        (1000, 250, 0): (10, 1354, 40, 5416), (1000, 500, 0): (10, 1354, 20, 2708),
        (5000, 1000, 271): (2, 271, 10, 1354), (5000, 1000, 270): (2, 270, 10, 1354),
        (5000, 500, 0): (2, 271, 20, 2708), (5000, 250, 270): (2, 270, 40, 5416),
    }.items():
        assert lib.gtp_cviirs_geometry(coarse, fine, width, *[ctypes.byref(v) for v in vals]) == 0
        assert tuple(v.value for v in vals) == want
    assert lib.gtp_cviirs_geometry(5000, 1000, 100, *[ctypes.byref(v) for v in vals]) == _capi.GTP_E_WIDTH_5KM
    assert b"less than 270 columns" in lib.gtp_last_error()
    assert lib.gtp_cviirs_geometry(2000, 250, 0, *[ctypes.byref(v) for v in vals]) == 1


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_capi, "_lib", None)
    monkeypatch.setattr(_capi, "LIB_PATH", "/nonexistent/libgtp_b200.so")
    with pytest.raises(ImportError, match="no CPU fallback"):
        _capi.lib()


# ---- argument validation (reference error types / messages) -----------
def test_required_keyword_arguments():
    a = np.zeros((10, 1354), np.float32)
    with pytest.raises(ValueError, match="'coarse_resolution' and 'fine_resolution' are required keyword arguments."):
        interpolate(a, a, a)


def test_expected_2d():
    a = np.zeros((10,), np.float32)
    with pytest.raises(ValueError, match="Expected 2D input arrays."):
        interpolate(a, a, a, coarse_resolution=1000, fine_resolution=250)


def test_5km_width_guard():
    a = np.zeros((2, 100), np.float32)
    with pytest.raises(NotImplementedError, match="Can't interpolate if 5km tiepoints have less than 270 columns."):
        modisinterpolator.modis_5km_to_1km(a, a, a)


def test_integer_dtype_is_a_type_error():
    a = np.zeros((10, 1354), np.int32)
    with pytest.raises(TypeError, match="No matching signature found"):
        modisinterpolator.modis_1km_to_250m(a, a, a)


def test_mixed_dtypes_is_a_value_error():
    a = np.zeros((10, 1354), np.float32)
    with pytest.raises(ValueError, match="Buffer dtype mismatch"):
        modisinterpolator.modis_1km_to_250m(a, a.astype(np.float64), a)


def test_mod03_argument_checks():
    """The mixed-type extension validates before it touches the device."""
    from geotiepoints_b200 import mod03
    lon = np.zeros((20, 1354), np.float32)
    zen = np.zeros((20, 1354), np.int16)
    with pytest.raises(TypeError):
        mod03.modis_1km_to_250m(lon.astype(np.float64), lon, zen)
    with pytest.raises(TypeError):
        mod03.modis_1km_to_250m(lon, lon, zen.astype(np.int32))
    with pytest.raises(TypeError):
        mod03.modis_1km_to_250m(lon, lon, zen, out_dtype=np.float16)
    with pytest.raises(ValueError):
        mod03.modis_1km_to_250m(lon[:15], lon[:15], zen[:15])
    with pytest.raises(ValueError):
        mod03.modis_1km_to_500m(lon[:, :1000], lon[:, :1000], zen[:, :1000])
    with pytest.raises(ValueError):
        mod03.modis_1km_to_250m(lon, lon[:10], zen)


def test_ecef_argument_checks_and_abi_validation():
    """The ECEF extension: host checks mirror the lon / lat operators; the C ABI validates its arguments
    before it looks for a device (no compute happens here)."""
    from geotiepoints_b200 import ecef, mod03
    a = np.zeros((15, 1354), np.float64)
    with pytest.raises(ValueError, match="whole scans"):
        ecef.modis_1km_to_250m(a, a, a)
    with pytest.raises(ValueError, match="whole scans"):
        ecef.simple_modis_1km_to_500m(a, a)
    with pytest.raises(NotImplementedError):
        ecef.interpolate(a[:2, :100], a[:2, :100], a[:2, :100], 5000, 1000, 100)
    with pytest.raises(TypeError):
        ecef.modis_1km_to_250m(a[:10].astype(np.int16), a[:10].astype(np.int16), a[:10].astype(np.int16))
    with pytest.raises(TypeError):
        mod03.modis_1km_to_250m(a[:10], a[:10], a[:10].astype(np.float32), ecef=True)
    lib = _capi.lib()
    buf = np.zeros(16, np.float64).ctypes.data
    assert lib.gtp_cviirs_interp_xyz_f64(buf, buf, buf, 1, 2000, 250, 0, buf, buf, buf, -1, None) == 1
    assert lib.gtp_cviirs_interp_xyz_f32(buf, buf, buf, 1, 5000, 1000, 100, buf, buf, buf, -1, None) == _capi.GTP_E_WIDTH_5KM
    assert lib.gtp_cviirs_interp_xyz_f64(buf, buf, buf, -1, 1000, 250, 0, buf, buf, buf, -1, None) == 4
    assert lib.gtp_cviirs_interp_xyz_f64(buf, buf, buf, 1, 1000, 250, 0, buf, buf, None, -1, None) == 3
    assert lib.gtp_cviirs_interp_xyz_f64(None, None, None, 0, 1000, 250, 0, None, None, None, -1, None) == 0
    assert lib.gtp_cviirs_interp_xyz_host_f64(buf, buf, buf, 1, 1000, 250, 0, buf, None, buf, -1) == 3
    assert lib.gtp_simple_interp_xyz_f64(buf, buf, 1, 1354, 3, buf, buf, buf, -1, None) == 1
    assert lib.gtp_simple_interp_xyz_host_f32(buf, buf, 1, 2, 4, buf, buf, buf, -1) == 4
    assert lib.gtp_cviirs_interp_xyz_mixed(buf, buf, buf, 0, 1.0, 0.0, 1, 5000, 1000, buf, buf, buf, 0, -1, None) == 1
    assert lib.gtp_cviirs_interp_xyz_mixed(buf, buf, buf, 7, 1.0, 0.0, 1, 1000, 250, buf, buf, buf, 0, -1, None) == 1
    assert lib.gtp_cviirs_interp_xyz_mixed_host(buf, buf, None, 0, 1.0, 0.0, 1, 1000, 250, buf, buf, buf, 0, -1) == 3
    assert lib.gtp_simple_interp_xyz_mixed(buf, buf, 1, 3, buf, buf, buf, 0, -1, None) == 1
    assert lib.gtp_simple_interp_xyz_mixed_host(buf, buf, 0, 4, buf, buf, buf, 1, -1) == 0


def test_sibling_abi_validation_before_any_device_work():
    """Legacy, VII and generic-grid entry points validate their arguments before they look for a device."""
    lib = _capi.lib()
    buf = np.zeros(64, np.float64).ctypes.data
    nan = float("nan")
    assert lib.gtp_legacy_interp_f64(buf, buf, 1, 7, buf, buf, -1, None) == 1
    assert lib.gtp_legacy_interp_host_f64(buf, buf, -1, 2, buf, buf, -1) == 4
    assert lib.gtp_vii_geo_interp_f64(buf, buf, 7, 4, 4, 8, 1, 0.8, buf, buf, -1, None) == 4          # 7 tie rows, 4 per scan
    assert lib.gtp_vii_interp_host_f64(buf, 8, 1, 4, 8, buf, -1) == 4                                  # one tie column
    grid = lambda *a: lib.gtp_grid_interp_f64(*a, -1, None)       # noqa: E731
    geo = lambda *a: lib.gtp_geogrid_interp_host_f64(*a, -1)      # noqa: E731
    assert grid(buf, 4, buf, 4, buf, 1, buf, 2, buf, 2, 2, 1, 0, nan, buf) == 1                        # scheme 2 does not exist
    assert grid(buf, 4, buf, 4, buf, 1, buf, 2, buf, 2, 1, 1, 3, nan, buf) == 1                        # out-of-range mode 3 neither
    assert grid(buf, 1, buf, 4, buf, 1, buf, 2, buf, 2, 1, 1, 0, nan, buf) == 4                        # one tie row
    assert grid(buf, 3, buf, 4, buf, 1, buf, 2, buf, 2, 3, 1, 0, nan, buf) == 4                        # cubic needs 4 points
    assert grid(buf, 4, buf, 4, buf, 0, buf, 2, buf, 2, 1, 1, 0, nan, buf) == 4                        # no arrays
    assert grid(buf, 4, buf, 4, buf, 1, buf, 2, buf, 2, 1, 1, 0, nan, None) == 3
    assert grid(buf, 4, buf, 4, buf, 1, buf, 0, buf, 2, 1, 1, 0, nan, None) == 0                       # nothing to do
    assert geo(buf, 4, buf, 4, buf, None, buf, 2, buf, 2, 1, 1, 0, nan, buf, buf) == 3
    assert geo(buf, 4, buf, 4, buf, buf, buf, 2, buf, 2, 0, 3, 2, nan, buf, None) == 3


def test_partial_scan_and_wrong_width_are_refused():
    a = np.zeros((15, 1354), np.float32)
    with pytest.raises(ValueError, match="whole scans"):
        modisinterpolator.modis_1km_to_250m(a, a, a)
    b = np.zeros((10, 1000), np.float32)
    with pytest.raises(ValueError, match="1354"):
        modisinterpolator.modis_1km_to_500m(b, b, b)
    with pytest.raises(ValueError, match="whole scans"):
        simple_modis_interpolator.modis_1km_to_250m(a, a)


def test_5km_poor_quality_warnings(monkeypatch):
    """modisinterpolator.py:45-47, :56-58 (tested in tests/test_modisinterpolator.py:130-134)."""
    import geotiepoints_b200.modisinterpolator as mi
    monkeypatch.setattr(mi, "interpolate", lambda *a, **k: ("lon", "lat"))
    a = np.zeros((2, 271), np.float32)
    for fn, expect in ((mi.modis_5km_to_500m, True), (mi.modis_5km_to_250m, True), (mi.modis_5km_to_1km, False),
                       (mi.modis_1km_to_250m, False)):
        with warnings.catch_warnings(record=True) as warns:
            warnings.simplefilter("always")
            fn(a, a, a)
        assert any("may result in poor quality" in str(w.message) for w in warns) == expect


def test_interpolator_geometry():
    m = MODISInterpolator(5000, 250, 270)
    assert (m._coarse_scan_length, m._coarse_scan_width, m._fine_pixels_per_coarse_pixel, m._fine_scan_width) == (2, 270, 20, 5416)
    m = MODISInterpolator(1000, 500)
    assert (m._coarse_scan_length, m._coarse_scan_width, m._fine_pixels_per_coarse_pixel, m._fine_scan_width) == (10, 1354, 2, 2708)
    assert [rows_per_scan_for_resolution(r) for r in (5000, 1000, 500, 250)] == [2, 10, 20, 40]


# ---- the dask adapter, against a fake dask ------------------------------------------
@pytest.fixture
def fake_da(monkeypatch):
    mod = types.SimpleNamespace(map_blocks=fake_dask.map_blocks, Array=fake_dask.Array)
    monkeypatch.setattr(_modis_utils, "da", mod)
    fake_dask.COMPUTES["n"] = 0
    fake_dask.BLOCK_SHAPES.clear()
    return mod


def _dummy_operator(factor):
    @scanline_mapblocks
    def op(lon, lat, coarse_resolution=0, fine_resolution=0):
        assert isinstance(lon, np.ndarray)
        out = np.repeat(np.repeat(lon, factor, axis=0), factor, axis=1)
        return out, out + 1
    return op


def test_dask_blocks_are_whole_scans_and_lazy(fake_da):
    lon = np.arange(60 * 1354, dtype=np.float32).reshape(60, 1354)
    d = fake_dask.from_array(lon, chunks=(25, 700))      # neither whole scans nor full width
    op = _dummy_operator(4)
    lons, lats = op(d, d, coarse_resolution=1000, fine_resolution=250)
    assert fake_dask.COMPUTES["n"] == 0                  # building the graph computes nothing
    assert lons.shape == (240, 5416) and lons.chunks == ((80, 80, 80), (5416,))
    out = lons.compute()
    assert fake_dask.BLOCK_SHAPES == [(20, 1354)] * 3     # 25 -> 20 rows = 2 whole scans
    np.testing.assert_array_equal(out, np.repeat(np.repeat(lon, 4, axis=0), 4, axis=1))
    np.testing.assert_array_equal(lats.compute(), out + 1)


def test_dask_partial_scan_is_refused(fake_da):
    lon = np.zeros((19, 1354), np.float32)
    d = fake_dask.from_array(lon, chunks=4096)
    with pytest.raises(ValueError, match=r"does not consist of whole scans \(10 rows per scan\)"):
        simple_modis_interpolator.modis_1km_to_250m(d, d)


def test_numpy_passthrough_does_not_touch_dask(fake_da):
    op = _dummy_operator(2)
    a = np.ones((10, 4), np.float32)
    lons, lats = op(a, a, coarse_resolution=1000, fine_resolution=500)
    assert isinstance(lons, np.ndarray) and lons.shape == (20, 8)


# ---- scan sharding ---------------------------------------------------------------------
@pytest.mark.parametrize("n_scans,world", [(288 * 203, 8), (203, 3), (7, 8), (0, 4), (58464, 1)])
def test_scan_ranges_partition(n_scans, world):
    ranges = [sharding.scan_range(r, world, n_scans) for r in range(world)]
    assert ranges[0][0] == 0 and ranges[-1][1] == n_scans
    for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [hi - lo for lo, hi in ranges]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.scan_range(world, world, n_scans)


# ---- zero-copy stacking of the operator outputs (what the dask adapter returns per block) ----
def test_stack_is_a_view_for_planes_of_one_block():
    from geotiepoints_b200 import _operator
    block = np.arange(2 * 6 * 8, dtype=np.float32).reshape(2, 6, 8)
    planes = (block[0], block[1])
    stacked = _operator.stack(planes)
    assert stacked.shape == (2, 6, 8) and np.shares_memory(stacked, block) and np.array_equal(stacked, block)
    three = np.zeros((3, 4, 4))
    assert np.shares_memory(_operator.stack(tuple(three[k] for k in range(3))), three)
    # anything else is copied, like the reference's np.concatenate (_modis_utils.pyx:191-198)
    a, b = np.ones((6, 8), np.float32), np.zeros((6, 8), np.float32)
    out = _operator.stack((a, b))
    assert out.shape == (2, 6, 8) and not np.shares_memory(out, a) and np.array_equal(out[0], a) and np.array_equal(out[1], b)
    swapped = _operator.stack((block[1], block[0]))          # same block, wrong order: not a view
    assert not np.shares_memory(swapped, block) and np.array_equal(swapped[0], block[1])
