"""ECEF output: the CUDA path, called through the C ABI, against the oracle's fine
Cartesian coordinates - the values the reference hands to ``xyz2lonlat``
(``_modis_interpolator.pyx:293``, ``_simple_modis_interpolator.pyx:63``).

Tolerances (tests/parity.py): float64 <= 1e-5 m (1e-10 deg of arc), float32 <= 1 float32 ulp, NaN
patterns identical.
"""
import ctypes

import numpy as np
import pytest

import gtp_oracle
from cases import CVIIRS_CONFIGS, DTYPES, INPUT_NAMES, SIMPLE_CONFIGS, cviirs_inputs
from geotiepoints_b200 import _capi, ecef, mod03, modisinterpolator, testing
from parity import assert_parity, assert_xyz_parity, compare_xyz

pytestmark = pytest.mark.gpu


def _oracle_xyz(a, coarse, fine, width):
    with np.errstate(all="ignore"):
        return gtp_oracle.cviirs_xyz(*a, coarse, fine, width, n_threads=8)


@pytest.mark.parametrize("dtype", DTYPES, ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("name", INPUT_NAMES)
def test_ecef_golden_inputs_all_configs(name, dtype, golden_inputs):
    for coarse, fine, width in CVIIRS_CONFIGS:
        a = cviirs_inputs(golden_inputs[name], coarse, width, dtype)
        got = ecef.interpolate(*a, coarse, fine, width)
        exp = _oracle_xyz(a, coarse, fine, width)
        assert len(got) == 3 and got[0].dtype == dtype and got[0].shape == exp[0].shape
        assert_xyz_parity(got, exp, what=f"{name} ecef {coarse}->{fine} w{width} {np.dtype(dtype).name}")


@pytest.mark.parametrize("dtype", DTYPES, ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("name", INPUT_NAMES)
def test_ecef_simple_golden_inputs(name, dtype, golden_inputs):
    lon1, lat1, _ = (a.astype(dtype) for a in golden_inputs[name])
    for coarse, fine in SIMPLE_CONFIGS:
        fn = ecef.simple_modis_1km_to_250m if fine == 250 else ecef.simple_modis_1km_to_500m
        got = fn(lon1, lat1)
        with np.errstate(all="ignore"):
            exp = gtp_oracle.simple_xyz(lon1, lat1, coarse, fine, n_threads=8)
        assert_xyz_parity(got, exp, what=f"{name} simple ecef {coarse}->{fine} {np.dtype(dtype).name}")


@pytest.mark.parametrize("dtype", DTYPES, ids=lambda d: np.dtype(d).name)
@pytest.mark.parametrize("granule", [0, 4, 14, 100, 287])
def test_ecef_synthetic_granules_1km(granule, dtype):
    lon, lat, satz = testing.modis_granule(granule, n_scans=12, dtype=dtype, quantize_satz=(granule % 2 == 1))
    for fn, fine in ((ecef.modis_1km_to_250m, 250), (ecef.modis_1km_to_500m, 500)):
        got = fn(lon, lat, satz)
        assert_xyz_parity(got, _oracle_xyz((lon, lat, satz), 1000, fine, 0),
                          what=f"granule {granule} ecef 1km->{fine} {np.dtype(dtype).name}")


def test_ecef_fast_and_generic_kernels_agree():
    """float64 at 1 km takes gtp_fast_xyz.cu; the rounding-faithful kernel is the tighter of the two."""
    import torch
    lib = _capi.lib()
    lon, lat, satz = testing.modis_granule(5, n_scans=8)
    dev = [torch.from_numpy(a).cuda() for a in (lon, lat, satz)]
    exp = _oracle_xyz((lon, lat, satz), 1000, 250, 0)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    stats = {}
    for name in ("gtp_cviirs_interp_xyz_f64", "gtp_cviirs_interp_xyz_generic_f64"):
        out = [torch.empty(exp[0].shape, dtype=torch.float64, device="cuda") for _ in range(3)]
        _capi.check(getattr(lib, name)(*[t.data_ptr() for t in dev], 8, 1000, 250, 0, *[o.data_ptr() for o in out], 0, st))
        torch.cuda.synchronize()
        stats[name] = assert_xyz_parity([o.cpu().numpy() for o in out], exp, what=name)
    # Same rounding points as the reference, and still ~2e-8 m from it: CUDA's and glibc's asin / sin
    # differ in the last bit, which the reference's c_expansion amplifies (tests/parity.py).  The fast
    # kernel is no further from the oracle than that.
    assert stats["gtp_cviirs_interp_xyz_generic_f64"]["max_abs_m"] < 1e-7, stats
    assert stats["gtp_cviirs_interp_xyz_f64"]["max_abs_m"] < 1e-7, stats
    # unaligned outputs fall back to the generic kernel instead of failing
    buf = [torch.empty(exp[0].size + 1, dtype=torch.float64, device="cuda") for _ in range(3)]
    out = [b[1:] for b in buf]
    _capi.check(lib.gtp_cviirs_interp_xyz_f64(*[t.data_ptr() for t in dev], 8, 1000, 250, 0, *[o.data_ptr() for o in out], 0, st))
    torch.cuda.synchronize()
    assert_xyz_parity([o.cpu().numpy().reshape(exp[0].shape) for o in out], exp, what="unaligned outputs")


def test_ecef_is_what_the_lonlat_path_back_transforms():
    """x, y, z -> lon, lat with the reference's formulas (_modis_utils.pyx:43-65) reproduces the lon / lat operator."""
    lon, lat, satz = testing.modis_granule(3, n_scans=6)
    x, y, z = ecef.modis_1km_to_250m(lon, lat, satz)
    exp = modisinterpolator.modis_1km_to_250m(lon, lat, satz)
    rho = np.sqrt(x * x + y * y)
    got_lon = np.rad2deg(np.arccos(x / rho)) * np.sign(y)
    low = np.abs(z) < 0.8 * 6370997.0
    got_lat = np.where(low, 90.0 - np.rad2deg(np.arccos(z / 6370997.0)),
                       np.sign(z) * (90.0 - np.rad2deg(np.arcsin(np.minimum(rho / 6370997.0, 1.0)))))
    assert_parity(got_lon, got_lat, *exp, what="ecef -> lon/lat vs lon/lat operator")


def test_ecef_torch_device_path_matches_host_path():
    import torch
    lon, lat, satz = testing.modis_granule(9, n_scans=5)
    host = ecef.modis_1km_to_250m(lon, lat, satz)
    dev = ecef.modis_1km_to_250m(*[torch.from_numpy(a).cuda() for a in (lon, lat, satz)])
    assert all(d.is_cuda for d in dev)
    for h, d in zip(host, dev):
        assert np.array_equal(h, d.cpu().numpy())
    lon5, lat5, satz5 = testing.subsample_5km(*testing.modis_granule(9, n_scans=10, dtype=np.float32), width=271)
    host = ecef.modis_5km_to_1km(lon5, lat5, satz5)
    dev = ecef.modis_5km_to_1km(*[torch.from_numpy(a).cuda() for a in (lon5, lat5, satz5)])
    for h, d in zip(host, dev):
        assert h.dtype == np.float32 and np.array_equal(h, d.cpu().numpy())


@pytest.mark.parametrize("fine", [250, 500])
@pytest.mark.parametrize("zen_kind", ["float32", "int16"])
@pytest.mark.parametrize("granule", [0, 14])
def test_mod03_ecef_matches_float64_reference(granule, zen_kind, fine):
    """MOD03 on-disk types in, float64 math: equals the oracle's float64 xyz on the widened inputs; the
    float32 output is that result rounded once."""
    import torch
    lon, lat, satz = testing.modis_granule(granule, n_scans=12, dtype=np.float32)
    if zen_kind == "int16":
        zen, scale = np.round(satz.astype(np.float64) / 0.01).astype(np.int16), 0.01
        satz64 = zen.astype(np.float64) * 0.01
    else:
        zen, scale, satz64 = satz, 1.0, satz.astype(np.float64)
    fn = mod03.modis_1km_to_250m if fine == 250 else mod03.modis_1km_to_500m
    exp = _oracle_xyz((lon.astype(np.float64), lat.astype(np.float64), satz64), 1000, fine, 0)
    got64 = fn(lon, lat, zen, zenith_scale=scale, out_dtype=np.float64, ecef=True)
    assert len(got64) == 3 and got64[0].dtype == np.float64
    assert_xyz_parity(got64, exp, what=f"mod03 ecef {zen_kind} granule {granule} ->{fine}")
    got32 = fn(lon, lat, zen, zenith_scale=scale, ecef=True)
    for g32, g64 in zip(got32, got64):
        assert g32.dtype == np.float32 and np.array_equal(g32, g64.astype(np.float32))
    tor = fn(*[torch.from_numpy(a).cuda() for a in (lon, lat, zen)], zenith_scale=scale, ecef=True)
    for t, g32 in zip(tor, got32):
        assert t.is_cuda and np.array_equal(t.cpu().numpy(), g32)
    # simple interpolator on the same lon / lat
    sfn = mod03.simple_modis_1km_to_250m if fine == 250 else mod03.simple_modis_1km_to_500m
    with np.errstate(all="ignore"):
        sexp = gtp_oracle.simple_xyz(lon.astype(np.float64), lat.astype(np.float64), 1000, fine, n_threads=8)
    assert_xyz_parity(sfn(lon, lat, out_dtype=np.float64, ecef=True), sexp, what=f"mod03 simple ecef granule {granule} ->{fine}")


def test_ecef_full_granule_sampled_scans():
    """A whole polar granule (203 scans) through the host pipeline; every 20th scan against the oracle."""
    lon, lat, satz = testing.modis_granule(4)
    got = ecef.modis_1km_to_250m(lon, lat, satz)
    worst = 0.0
    for s in range(0, 203, 20):
        a = [np.ascontiguousarray(v[s * 10:(s + 1) * 10]) for v in (lon, lat, satz)]
        stats = assert_xyz_parity([g[s * 40:(s + 1) * 40] for g in got], _oracle_xyz(a, 1000, 250, 0), what=f"scan {s}")
        worst = max(worst, stats["max_abs_m"])
    assert worst < 1e-5


@pytest.mark.parametrize("config", ["f64_250m", "f64_500m", "mixed_f32_250m", "simple_f64_250m"])
def test_ecef_kernels_stay_inside_their_output_buffers(config):
    import torch
    lib = _capi.lib()
    lon, lat, satz = (np.ascontiguousarray(a[-40:]) for a in testing.modis_granule(4))
    pad, sentinel, n_scans = 4096, -12345.0, 4
    mixed = config.startswith("mixed")
    npdt, tdt = (np.float32, torch.float32) if mixed else (np.float64, torch.float64)
    p = 2 if config.endswith("500m") else 4
    dev = [torch.from_numpy(a.astype(npdt)).cuda() for a in (lon, lat, satz)]
    n_out = 40 * p * 1354 * p
    bufs = [torch.full((n_out + 2 * pad,), sentinel, dtype=tdt, device="cuda") for _ in range(3)]
    optr = [b[pad:pad + n_out].data_ptr() for b in bufs]
    ptr = [t.data_ptr() for t in dev]
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    if config.startswith("f64"):
        rc = lib.gtp_cviirs_interp_xyz_f64(*ptr, n_scans, 1000, 1000 // p, 0, *optr, 0, st)
    elif mixed:
        rc = lib.gtp_cviirs_interp_xyz_mixed(*ptr, 0, 1.0, 0.0, n_scans, 1000, 250, *optr, 0, 0, st)
    else:
        rc = lib.gtp_simple_interp_xyz_f64(ptr[0], ptr[1], n_scans, 1354, 4, *optr, 0, st)
    _capi.check(rc)
    torch.cuda.synchronize()
    for b in bufs:
        assert bool((b[:pad] == sentinel).all()) and bool((b[pad + n_out:] == sentinel).all()), "wrote outside the output"
        assert not bool((b[pad:pad + n_out] == sentinel).any()), "left output pixels unwritten"


def test_ecef_argument_errors():
    lon, lat, satz = testing.modis_granule(0, n_scans=1, dtype=np.float32)
    with pytest.raises(ValueError):
        mod03.modis_1km_to_250m(lon[:, :100], lat[:, :100], satz[:, :100], ecef=True)
    with pytest.raises(NotImplementedError):
        ecef.interpolate(lon[:2, :200], lat[:2, :200], satz[:2, :200], 5000, 1000, 200)
    stats = compare_xyz(ecef.modis_1km_to_250m(lon, lat, satz), _oracle_xyz((lon, lat, satz), 1000, 250, 0))
    assert stats["max_ulps"] <= 1.0
