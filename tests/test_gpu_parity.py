"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle.

Every case runs the same inputs through ``libgtp_b200.so`` (host entry points for numpy,
device entry points for torch CUDA tensors) and through ``oracle/libgtp_oracle.so`` and
compares with the tolerances of BASELINE.json (tests/parity.py: 1e-10 deg float64,
1e-5 deg / 1 ulp float32, conditioning band reported separately).
"""
import numpy as np
import pytest

import gtp_oracle
from cases import CVIIRS_CONFIGS, DTYPES, INPUT_NAMES, SIMPLE_CONFIGS, cviirs_inputs
from geotiepoints_b200 import _capi, modisinterpolator, simple_modis_interpolator, testing
from geotiepoints_b200._modis_interpolator import interpolate
from parity import assert_parity, compare

pytestmark = pytest.mark.gpu


def _oracle_cviirs(a, coarse, fine, width):
    with np.errstate(all="ignore"):
        return gtp_oracle.cviirs(*a, coarse, fine, width, n_threads=8)


@pytest.mark.parametrize("dtype", DTYPES, ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("name", INPUT_NAMES)
def test_cviirs_golden_inputs_all_configs(name, dtype, golden_inputs):
    for coarse, fine, width in CVIIRS_CONFIGS:
        a = cviirs_inputs(golden_inputs[name], coarse, width, dtype)
        got = interpolate(*a, coarse_resolution=coarse, fine_resolution=fine, coarse_scan_width=width)
        exp = _oracle_cviirs(a, coarse, fine, width)
        assert got[0].dtype == dtype and got[0].shape == exp[0].shape
        assert_parity(*got, *exp, what=f"{name} cviirs {coarse}->{fine} w{width} {np.dtype(dtype).name}")


@pytest.mark.parametrize("dtype", DTYPES, ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("name", INPUT_NAMES)
def test_simple_golden_inputs(name, dtype, golden_inputs):
    lon1, lat1, _ = (a.astype(dtype) for a in golden_inputs[name])
    for coarse, fine in SIMPLE_CONFIGS:
        got = simple_modis_interpolator.interpolate_geolocation_cartesian(
            lon1, lat1, coarse_resolution=coarse, fine_resolution=fine)
        with np.errstate(all="ignore"):
            exp = gtp_oracle.simple(lon1, lat1, coarse, fine, n_threads=8)
        assert_parity(*got, *exp, what=f"{name} simple {coarse}->{fine} {np.dtype(dtype).name}")


@pytest.mark.parametrize("dtype", DTYPES, ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("granule", [0, 5, 9, 14, 100, 287])
def test_cviirs_synthetic_granules_1km(granule, dtype):
    """Orbit-synthetic granules over the whole day: equator, both polar caps, antimeridian."""
    lon, lat, satz = testing.modis_granule(granule, n_scans=12, dtype=dtype, quantize_satz=(granule % 2 == 1))
    for fine in (250, 500):
        got = interpolate(lon, lat, satz, coarse_resolution=1000, fine_resolution=fine)
        exp = _oracle_cviirs((lon, lat, satz), 1000, fine, 0)
        assert_parity(*got, *exp, what=f"granule {granule} 1km->{fine} {np.dtype(dtype).name}")


@pytest.mark.parametrize("granule", [4, 14, 24, 34])
def test_polar_granules_sampled_scans(granule):
    """Scans spread over the granules that sweep over a pole (the first 12 scans of a granule never get
    there): long polar series, wrap-around quads and the rounding-faithful fallback on the GPU, for the
    CVIIRS and the simple interpolator."""
    lon, lat, satz = testing.modis_granule(granule)
    picks = np.concatenate([np.arange(s * 10, s * 10 + 10) for s in range(0, 203, 17)])
    lon, lat, satz = (np.ascontiguousarray(a[picks]) for a in (lon, lat, satz))
    assert np.abs(lat).max() > 88.0
    for fine in (250, 500):
        got = interpolate(lon, lat, satz, coarse_resolution=1000, fine_resolution=fine)
        exp = _oracle_cviirs((lon, lat, satz), 1000, fine, 0)
        assert_parity(*got, *exp, what=f"polar granule {granule} 1km->{fine}")
        got = simple_modis_interpolator.interpolate_geolocation_cartesian(
            lon, lat, coarse_resolution=1000, fine_resolution=fine)
        with np.errstate(all="ignore"):
            exp = gtp_oracle.simple(lon, lat, 1000, fine, n_threads=8)
        assert_parity(*got, *exp, what=f"polar granule {granule} simple 1km->{fine}")


@pytest.mark.parametrize("dtype", DTYPES, ids=lambda d: np.dtype(d).name)
def test_cviirs_full_granule_5km(dtype):
    """BASELINE config 3: 406 x 271|270 tie points -> 2030 x 1354."""
    lon, lat, satz = testing.modis_granule(33, dtype=dtype)
    for width in (271, 270):
        a = testing.subsample_5km(lon, lat, satz, width)
        assert a[0].shape == (406, width)
        got = modisinterpolator.modis_5km_to_1km(*a)
        exp = _oracle_cviirs(a, 5000, 1000, width)
        assert got[0].shape == (2030, 1354)
        assert_parity(*got, *exp, what=f"5km({width})->1km {np.dtype(dtype).name}")


@pytest.mark.parametrize("dtype", DTYPES, ids=lambda d: np.dtype(d).name)
def test_cviirs_full_granule_1km_to_250m(dtype):
    """BASELINE config 2 at full size (oracle on 8 threads finishes in a few seconds)."""
    lon, lat, satz = testing.modis_granule(141, dtype=dtype)
    got = modisinterpolator.modis_1km_to_250m(lon, lat, satz)
    exp = _oracle_cviirs((lon, lat, satz), 1000, 250, 0)
    assert got[0].shape == (8120, 5416)
    s = assert_parity(*got, *exp, what=f"full granule 1km->250m {np.dtype(dtype).name}")
    print("\nparity full granule", np.dtype(dtype).name, s)


@pytest.mark.parametrize("dtype", DTYPES, ids=lambda d: np.dtype(d).name)
def test_simple_full_granule_and_odd_width(dtype):
    lon, lat, _ = testing.modis_granule(77, n_scans=40, dtype=dtype)
    for fine in (250, 500):
        got = simple_modis_interpolator.modis_1km_to_250m(lon, lat) if fine == 250 else \
            simple_modis_interpolator.modis_1km_to_500m(lon, lat)
        exp = gtp_oracle.simple(lon, lat, 1000, fine, n_threads=8)
        assert_parity(*got, *exp, what=f"simple 1km->{fine} {np.dtype(dtype).name}")
    a = np.ascontiguousarray(lon[:20, :131]), np.ascontiguousarray(lat[:20, :131])
    got = simple_modis_interpolator.interpolate_geolocation_cartesian(*a, coarse_resolution=1000, fine_resolution=250)
    exp = gtp_oracle.simple(*a, 1000, 250)
    assert_parity(*got, *exp, what="simple width 131")


def test_generic_f64_kernel_matches_oracle_tightly():
    """The rounding-faithful float64 kernel differs from the reference only through CUDA's
    libm (<= 2 ulp) - checked at a tolerance 100x tighter than the contract, outside the band."""
    import ctypes
    import torch
    lon, lat, satz = testing.modis_granule(20, n_scans=6)
    exp = _oracle_cviirs((lon, lat, satz), 1000, 250, 0)
    d = [torch.from_numpy(a).cuda() for a in (lon, lat, satz)]
    out_lon = torch.empty(exp[0].shape, dtype=torch.float64, device="cuda")
    out_lat = torch.empty_like(out_lon)
    rc = _capi.lib().gtp_cviirs_interp_generic_f64(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), 6, 1000, 250, 0,
                                                    out_lon.data_ptr(), out_lat.data_ptr(), 0,
                                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _capi.check(rc)
    torch.cuda.synchronize()
    s = compare(out_lon.cpu().numpy(), out_lat.cpu().numpy(), *exp)
    assert s["max_dlat"] < 1e-12 and s["max_dlon_outside_band"] < 1e-12, s


def test_generic_f64_simple_kernel_matches_oracle_tightly():
    """The rounding-faithful simple kernel (map_coordinates + extrapolation, tap by tap)."""
    import ctypes
    import torch
    lon, lat, _ = testing.modis_granule(21, n_scans=6)
    exp = gtp_oracle.simple(lon, lat, 1000, 250, n_threads=8)
    d = [torch.from_numpy(a).cuda() for a in (lon, lat)]
    out_lon = torch.empty(exp[0].shape, dtype=torch.float64, device="cuda")
    out_lat = torch.empty_like(out_lon)
    rc = _capi.lib().gtp_simple_interp_generic_f64(d[0].data_ptr(), d[1].data_ptr(), 6, 1354, 4,
                                                    out_lon.data_ptr(), out_lat.data_ptr(), 0,
                                                    ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    _capi.check(rc)
    torch.cuda.synchronize()
    s = compare(out_lon.cpu().numpy(), out_lat.cpu().numpy(), *exp)
    assert s["max_dlat"] < 1e-12 and s["max_dlon_outside_band"] < 1e-12, s


def test_torch_device_path_matches_host_path():
    import torch
    lon, lat, satz = testing.modis_granule(3, n_scans=5)
    host = modisinterpolator.modis_1km_to_250m(lon, lat, satz)
    dev = modisinterpolator.modis_1km_to_250m(*[torch.from_numpy(a).cuda() for a in (lon, lat, satz)])
    assert dev[0].is_cuda and dev[0].dtype == torch.float64
    np.testing.assert_array_equal(dev[0].cpu().numpy(), host[0])
    np.testing.assert_array_equal(dev[1].cpu().numpy(), host[1])
    host = simple_modis_interpolator.modis_1km_to_500m(lon.astype(np.float32), lat.astype(np.float32))
    dev = simple_modis_interpolator.modis_1km_to_500m(*[torch.from_numpy(a.astype(np.float32)).cuda() for a in (lon, lat)])
    np.testing.assert_array_equal(dev[0].cpu().numpy(), host[0])
    np.testing.assert_array_equal(dev[1].cpu().numpy(), host[1])


def test_empty_and_strided_inputs():
    lon, lat, satz = testing.modis_granule(3, n_scans=2, dtype=np.float32)
    got = modisinterpolator.modis_1km_to_500m(lon[:0], lat[:0], satz[:0])
    assert got[0].shape == (0, 2708) and got[1].shape == (0, 2708)
    # non-contiguous views are made contiguous, as the reference does (:136-138)
    big = np.zeros((20, 2708), np.float32)
    big[:, ::2] = lon
    ref = modisinterpolator.modis_1km_to_500m(lon, lat, satz)
    got = modisinterpolator.modis_1km_to_500m(big[:, ::2], lat, satz)
    np.testing.assert_array_equal(got[0], ref[0])
    # inputs are never modified
    lon2 = lon.copy()
    modisinterpolator.modis_1km_to_250m(lon, lat, satz)
    np.testing.assert_array_equal(lon, lon2)


def test_launch_counter_moves():
    lon, lat, satz = testing.modis_granule(3, n_scans=1, dtype=np.float32)
    before = _capi.launch_count()
    modisinterpolator.modis_1km_to_250m(lon, lat, satz)
    assert _capi.launch_count() == before + 1


# ---- mixed-type extension (mod03): on-disk types in, float64 math ---------------------------------
def _mod03_inputs(granule, n_scans, zen_kind):
    lon, lat, satz = testing.modis_granule(granule, n_scans=n_scans, dtype=np.float32)
    if zen_kind == "int16":
        zen = np.round(satz.astype(np.float64) / 0.01).astype(np.int16)
        return lon, lat, zen, 0.01, zen.astype(np.float64) * 0.01
    return lon, lat, satz, 1.0, satz.astype(np.float64)


@pytest.mark.parametrize("fine", [250, 500])
@pytest.mark.parametrize("zen_kind", ["float32", "int16"])
@pytest.mark.parametrize("granule", [0, 5, 14])
def test_mod03_mixed_types_match_float64_reference(granule, zen_kind, fine):
    """float32 lon/lat + float32|int16 zenith -> float64 math: equals the oracle's float64 path on the
    widened inputs (1e-10 deg for float64 output; the float32 output is that result rounded once)."""
    import torch
    from geotiepoints_b200 import mod03
    lon, lat, zen, scale, satz64 = _mod03_inputs(granule, 12, zen_kind)
    fn = mod03.modis_1km_to_250m if fine == 250 else mod03.modis_1km_to_500m
    exp = _oracle_cviirs((lon.astype(np.float64), lat.astype(np.float64), satz64), 1000, fine, 0)
    got64 = fn(lon, lat, zen, zenith_scale=scale, out_dtype=np.float64)
    assert got64[0].dtype == np.float64 and got64[0].shape == exp[0].shape
    assert_parity(*got64, *exp, what=f"mod03 {zen_kind} granule {granule} ->{fine} float64 out")
    got32 = fn(lon, lat, zen, zenith_scale=scale)
    assert got32[0].dtype == np.float32
    for g32, g64, name in ((got32[0], got64[0], "lon"), (got32[1], got64[1], "lat")):
        assert np.array_equal(g32, g64.astype(np.float32)), f"float32 {name} is not the rounded float64 result"
    # torch CUDA tensors: zero copy, same numbers
    dev = [torch.from_numpy(a).cuda() for a in (lon, lat, zen)]
    tor = fn(*dev, zenith_scale=scale)
    assert tor[0].is_cuda and tor[0].dtype == torch.float32
    assert np.array_equal(tor[0].cpu().numpy(), got32[0]) and np.array_equal(tor[1].cpu().numpy(), got32[1])


@pytest.mark.parametrize("width", [271, 270])
@pytest.mark.parametrize("zen_kind", ["float32", "int16"])
@pytest.mark.parametrize("granule", [0, 14])
def test_mod03_5km_matches_float64_reference(granule, zen_kind, width):
    from geotiepoints_b200 import mod03
    lon, lat, zen, scale, satz64 = _mod03_inputs(granule, 203, zen_kind)
    lon5, lat5, zen5 = testing.subsample_5km(lon, lat, zen, width)
    _, _, satz5 = testing.subsample_5km(lon, lat, satz64, width)
    exp = _oracle_cviirs((lon5.astype(np.float64), lat5.astype(np.float64), satz5), 5000, 1000, width)
    got64 = mod03.modis_5km_to_1km(lon5, lat5, zen5, zenith_scale=scale, out_dtype=np.float64)
    assert_parity(*got64, *exp, what=f"mod03 5km({width}) {zen_kind} granule {granule} float64 out")
    got32 = mod03.modis_5km_to_1km(lon5, lat5, zen5, zenith_scale=scale)
    assert got32[0].dtype == np.float32 and got32[0].shape == (2030, 1354)
    assert np.array_equal(got32[0], got64[0].astype(np.float32)) and np.array_equal(got32[1], got64[1].astype(np.float32))


@pytest.mark.parametrize("fine", [250, 500])
@pytest.mark.parametrize("granule", [0, 5, 14])
def test_mod03_simple_matches_float64_reference(granule, fine):
    from geotiepoints_b200 import mod03
    lon, lat, _ = testing.modis_granule(granule, n_scans=12, dtype=np.float32)
    fn = mod03.simple_modis_1km_to_250m if fine == 250 else mod03.simple_modis_1km_to_500m
    with np.errstate(all="ignore"):
        exp = gtp_oracle.simple(lon.astype(np.float64), lat.astype(np.float64), 1000, fine, n_threads=8)
    got64 = fn(lon, lat, out_dtype=np.float64)
    assert_parity(*got64, *exp, what=f"mod03 simple granule {granule} ->{fine} float64 out")
    got32 = fn(lon, lat)
    assert got32[0].dtype == np.float32
    assert np.array_equal(got32[0], got64[0].astype(np.float32)) and np.array_equal(got32[1], got64[1].astype(np.float32))


def test_concurrent_host_calls_from_threads():
    """dask's threaded scheduler calls the operator from several threads at once: every call checks out
    its own streams + scratch (csrc/gtp_capi.cu HostPipe), results must equal the serial ones bit for bit."""
    from concurrent.futures import ThreadPoolExecutor
    inputs = [testing.modis_granule(g, n_scans=30 + 7 * g) for g in range(6)]
    serial = [modisinterpolator.modis_1km_to_250m(*a) for a in inputs]
    serial = [(np.array(lo), np.array(la)) for lo, la in serial]
    with ThreadPoolExecutor(6) as pool:
        for _ in range(3):
            got = list(pool.map(lambda a: modisinterpolator.modis_1km_to_250m(*a), inputs))
            for (lo, la), (slo, sla) in zip(got, serial):
                assert np.array_equal(lo, slo, equal_nan=True) and np.array_equal(la, sla, equal_nan=True)
    assert _capi.lib().gtp_trim() == 0


@pytest.mark.parametrize("config", ["1km_250m", "1km_500m", "5km_271", "5km_270", "simple_250m", "mixed_250m_f32", "mixed_5km_f32"])
def test_kernels_stay_inside_their_output_buffers(config):
    """compute-sanitizer is not available on the GPU pool, so: outputs are carved out of the middle of a
    sentinel-filled allocation; after the launch everything outside them must still be the sentinel and
    everything inside must have been written.  A polar granule tail exercises every evaluation mode."""
    import ctypes
    import torch
    lib = _capi.lib()
    lon, lat, satz = (np.ascontiguousarray(a[-40:]) for a in testing.modis_granule(4))
    pad, sentinel = 4096, -12345.0
    tdt, npdt, out_dt = torch.float64, np.float64, torch.float64
    if config.startswith("mixed"):
        tdt, npdt, out_dt = torch.float32, np.float32, torch.float32
    arrs = [a.astype(npdt) for a in (lon, lat, satz)]
    n_scans, out_shape = 4, None
    if "5km" in config:
        width = 270 if config.endswith("270") else 271
        arrs = list(testing.subsample_5km(*arrs, width=width))
        n_scans, out_shape = arrs[0].shape[0] // 2, (arrs[0].shape[0] * 5, 1354)
    else:
        p = 2 if "500m" in config else 4
        out_shape = (arrs[0].shape[0] * p, 1354 * p)
    dev = [torch.from_numpy(a).cuda() for a in arrs]
    n_out = out_shape[0] * out_shape[1]
    bufs = [torch.full((n_out + 2 * pad,), sentinel, dtype=out_dt, device="cuda") for _ in range(2)]
    outs = [b[pad:pad + n_out] for b in bufs]
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ptr = [t.data_ptr() for t in dev]
    optr = [o.data_ptr() for o in outs]
    if config in ("1km_250m", "1km_500m"):
        rc = lib.gtp_cviirs_interp_f64(*ptr, n_scans, 1000, 250 if config == "1km_250m" else 500, 0, *optr, 0, st)
    elif config.startswith("5km"):
        rc = lib.gtp_cviirs_interp_f64(*ptr, n_scans, 5000, 1000, arrs[0].shape[1], *optr, 0, st)
    elif config == "simple_250m":
        rc = lib.gtp_simple_interp_f64(ptr[0], ptr[1], n_scans, 1354, 4, *optr, 0, st)
    elif config == "mixed_250m_f32":
        rc = lib.gtp_cviirs_interp_mixed(*ptr, 0, 1.0, 0.0, n_scans, 1000, 250, 0, *optr, 0, 0, st)
    else:
        sub = testing.subsample_5km(*arrs, width=271)
        dev = [torch.from_numpy(a).cuda() for a in sub]
        n_scans = sub[0].shape[0] // 2
        n_out = n_scans * 10 * 1354
        outs = [b[pad:pad + n_out] for b in bufs]
        rc = lib.gtp_cviirs_interp_mixed(*[t.data_ptr() for t in dev], 0, 1.0, 0.0, n_scans, 5000, 1000, 271,
                                         *[o.data_ptr() for o in outs], 0, 0, st)
    _capi.check(rc)
    torch.cuda.synchronize()
    for b in bufs:
        assert bool((b[:pad] == sentinel).all()) and bool((b[pad + n_out:] == sentinel).all()), "wrote outside the output"
        assert not bool((b[pad:pad + n_out] == sentinel).any()), "left output pixels unwritten"


def _assert_fill_parity(got, exp, rows_per_scan, what):
    """Scans 0 (valid) and 1 (all fill) to the usual tolerances; quads that mix valid and fill corners are
    chords deep inside the sphere where the reference's own formulas are ill-conditioned: 1e-7 deg."""
    n = 2 * rows_per_scan
    assert_parity(got[0][:n], got[1][:n], exp[0][:n], exp[1][:n], what=what + " (scans 0, 1)")
    s = compare(got[0], got[1], *exp)
    assert s["nan_pattern_equal"], (what, s)
    if got[0].dtype == np.float64:
        assert s["max_dlat"] < 1e-7 and max(s["max_dlon_outside_band"], s["max_dlon_in_band"]) < 1e-7, (what, s)
    else:
        assert s["n_over_tol_and_ulp"] == 0, (what, s)


@pytest.mark.parametrize("dtype", DTYPES, ids=lambda d: np.dtype(d).name)
def test_fill_values_fold_like_the_reference(dtype):
    """-999 (MOD03 _FillValue) in lon / lat comes back folded (81 deg), as the reference's
    lonlat2xyz -> xyz2lonlat does, on every fast path; int16 zenith fill (-32767) through the mod03 entry."""
    lon, lat, satz = testing.fill_value_granule(dtype=dtype)
    name = np.dtype(dtype).name
    for fine in (250, 500):
        exp = _oracle_cviirs((lon, lat, satz), 1000, fine, 0)
        assert np.nanmax(np.abs(exp[1])) <= 90.0 and np.nanmax(np.abs(exp[0])) <= 180.0
        got = interpolate(lon, lat, satz, coarse_resolution=1000, fine_resolution=fine)
        _assert_fill_parity(got, exp, 10000 // fine, f"fill values cviirs 1km->{fine} {name}")
        with np.errstate(all="ignore"):
            exp_s = gtp_oracle.simple(lon, lat, 1000, fine)
        got_s = simple_modis_interpolator.interpolate_geolocation_cartesian(lon, lat, coarse_resolution=1000, fine_resolution=fine)
        _assert_fill_parity(got_s, exp_s, 10000 // fine, f"fill values simple 1km->{fine} {name}")
    for width in (271, 270):
        a5 = testing.subsample_5km(lon, lat, satz, width)
        exp5 = _oracle_cviirs(a5, 5000, 1000, width)
        got5 = interpolate(*a5, coarse_resolution=5000, fine_resolution=1000, coarse_scan_width=width)
        _assert_fill_parity(got5, exp5, 10, f"fill values cviirs 5km({width})->1km {name}")


def test_fill_values_mod03_types():
    from geotiepoints_b200 import mod03
    lon, lat, satz = testing.fill_value_granule(dtype=np.float32)
    zen = np.round(satz.astype(np.float64) / 0.01).astype(np.int16)
    zen[10:20] = -32767                                    # the SensorZenith fill value, taken as it is
    satz64 = zen.astype(np.float64) * 0.01
    exp = _oracle_cviirs((lon.astype(np.float64), lat.astype(np.float64), satz64), 1000, 250, 0)
    got = mod03.modis_1km_to_250m(lon, lat, zen, zenith_scale=0.01, out_dtype=np.float64)
    _assert_fill_parity(got, exp, 40, "fill values mod03 f32 + i16 -> float64")


@pytest.mark.parametrize("name", ["prime_meridian", "antimeridian", "pole", "south_pole"])
def test_band_is_the_references_noise_not_the_series(name, golden_inputs):
    """VERDICT r1, "what's weak" 1: inside the conditioning bands (|sin lon| < 2e-4, ~4 km around a pole) the
    float64 tolerance is a rounding-noise model of the reference's own acos(x / sqrt(x^2 + y^2)).  Shown here:
    (1) the ROUNDING-FAITHFUL kernel (the reference's formulas operation by operation, CUDA libm instead of
    glibc) lands inside the same model; (2) next to the 0 / 180 deg meridians the deviation of the series from
    the oracle IS the oracle's error: against atan2(y, x) of the oracle's own Cartesian coordinates evaluated
    in 80-bit arithmetic the series holds the stated 1e-10 deg, the oracle (= the reference) does not - its
    acos argument x / rho is quantised to 1.1e-16 next to +-1, i.e. to steps of 1.1e-16 / |sin lon| rad."""
    import ctypes
    import torch
    from parity import BAND, lon_diff
    lib = _capi.lib()
    lon, lat, satz = (np.ascontiguousarray(a.astype(np.float64)) for a in golden_inputs[name])
    n_scans = lon.shape[0] // 10
    dev = [torch.from_numpy(a).cuda() for a in (lon, lat, satz)]
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    exp = _oracle_cviirs((lon, lat, satz), 1000, 250, 0)
    stats, got = {}, {}
    for kernel in ("gtp_cviirs_interp_generic_f64", "gtp_cviirs_interp_f64"):
        outs = [torch.empty((n_scans * 40, 5416), dtype=torch.float64, device="cuda") for _ in range(2)]
        _capi.check(getattr(lib, kernel)(*[t.data_ptr() for t in dev], n_scans, 1000, 250, 0, *[o.data_ptr() for o in outs], 0, st))
        torch.cuda.synchronize()
        got[kernel] = outs[0].cpu().numpy()
        stats[kernel] = assert_parity(got[kernel], outs[1].cpu().numpy(), *exp, what=f"{kernel} {name}")
    gen, fast = stats["gtp_cviirs_interp_generic_f64"], stats["gtp_cviirs_interp_f64"]
    assert gen["n_in_band"] == fast["n_in_band"] > 0, (gen, fast)
    assert gen["n_in_band_over_conditioning_bound"] == 0 and fast["n_in_band_over_conditioning_bound"] == 0
    print(f"{name}: in-band pixels {gen['n_in_band']} of {gen['n']} ({gen['n_in_band'] / gen['n']:.2e}); max |dlon| vs oracle in "
          f"band: faithful kernel {gen['max_dlon_in_band']:.2e}, series {fast['max_dlon_in_band']:.2e} deg")
    if np.finfo(np.longdouble).nmant < 63 or "meridian" not in name:
        return
    with np.errstate(all="ignore"):
        x, y, _ = (a.astype(np.longdouble) for a in gtp_oracle.cviirs_xyz(lon, lat, satz, 1000, 250))
    truth = np.degrees(np.arctan2(y, x))
    band = np.abs(np.sin(np.deg2rad(exp[0]))) < BAND
    err = {k: float(np.max(np.where(band, lon_diff(v.astype(np.longdouble), truth).astype(np.float64), 0.0)))
           for k, v in (("oracle", exp[0]), ("faithful", got["gtp_cviirs_interp_generic_f64"]), ("series", got["gtp_cviirs_interp_f64"]))}
    print(f"{name}: max |lon - atan2(y, x) in 80-bit| inside the band: {err}")
    assert err["series"] <= 1e-10, err                       # the fast path meets the stated tolerance against the truth
    assert err["oracle"] >= 0.5 * fast["max_dlon_in_band"], (err, fast)        # what it differs from the oracle by is the oracle's error
