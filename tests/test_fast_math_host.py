"""The float64 fast-path math of the CUDA kernel, run on the CPU (tests/host_emulation.cpp
includes csrc/gtp_fast_math.cuh) and checked against the oracle.  This pins the series, the
acceptance test (`Quad::mode`) and the fallback without a GPU; the GPU tests then only have to
show that the kernel's warp orchestration feeds the same math."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import gtp_oracle
from conftest import ROOT
from geotiepoints_b200 import testing
from parity import assert_parity

SRC = os.path.join(ROOT, "tests", "host_emulation.cpp")
LIB = os.path.join(ROOT, "tests", "libgtp_host_emulation.so")
HDRS = [os.path.join(ROOT, "python-geotiepoints_b200", "geotiepoints_b200", "csrc", n)
        for n in ("gtp_fast_math.cuh", "gtp_fast_f32_math.cuh", "gtp_common.cuh")]


@pytest.fixture(scope="module")
def emul():
    if not os.path.exists(LIB) or any(os.path.getmtime(LIB) < os.path.getmtime(f) for f in [SRC] + HDRS):
        # -ffp-contract=off: only the explicit fma() calls of the series are fused, as on the GPU
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-mfma",
                               SRC, "-o", LIB])
    lib = ctypes.CDLL(LIB)
    dp = ctypes.POINTER(ctypes.c_double)

    def run(lon, lat, satz, fine):
        ns, p = lon.shape[0] // 10, 1000 // fine
        out_lon = np.full((ns * 10 * p, 1354 * p), np.nan)
        out_lat = np.full_like(out_lon, np.nan)
        modes = (ctypes.c_long * 6)()
        rc = lib.gtp_emul_cviirs_fast_f64(lon.ctypes.data_as(dp), lat.ctypes.data_as(dp), satz.ctypes.data_as(dp),
                                          ctypes.c_long(ns), fine, out_lon.ctypes.data_as(dp),
                                          out_lat.ctypes.data_as(dp), modes)
        assert rc == 0
        return out_lon, out_lat, dict(zip(("slow", "low", "high", "both", "small", "wide"), modes))

    def run_simple(lon, lat, fine):
        ns, p = lon.shape[0] // 10, 1000 // fine
        out_lon = np.full((ns * 10 * p, 1354 * p), np.nan)
        out_lat = np.full_like(out_lon, np.nan)
        modes = (ctypes.c_long * 6)()
        rc = lib.gtp_emul_simple_fast_f64(lon.ctypes.data_as(dp), lat.ctypes.data_as(dp), ctypes.c_long(ns), fine,
                                          out_lon.ctypes.data_as(dp), out_lat.ctypes.data_as(dp), modes)
        assert rc == 0
        return out_lon, out_lat, dict(zip(("slow", "low", "high", "both", "small", "wide"), modes))
    run.simple = run_simple

    def run_5km(lon, lat, satz):
        ns, w = lon.shape[0] // 2, lon.shape[1]
        out_lon = np.full((ns * 10, 1354), np.nan)
        out_lat = np.full_like(out_lon, np.nan)
        modes = (ctypes.c_long * 6)()
        rc = lib.gtp_emul_cviirs_fast5_f64(lon.ctypes.data_as(dp), lat.ctypes.data_as(dp), satz.ctypes.data_as(dp),
                                           ctypes.c_long(ns), w, out_lon.ctypes.data_as(dp),
                                           out_lat.ctypes.data_as(dp), modes)
        assert rc == 0
        return out_lon, out_lat, dict(zip(("slow", "low", "high", "both", "small", "wide"), modes))
    run.five_km = run_5km
    fp = ctypes.POINTER(ctypes.c_float)

    def run_f32(lon, lat, satz, fine):
        ns, p = lon.shape[0] // 10, 1000 // fine
        out_lon = np.full((ns * 10 * p, 1354 * p), np.nan, np.float32)
        out_lat = np.full_like(out_lon, np.nan)
        modes = (ctypes.c_long * 8)()
        rc = lib.gtp_emul_cviirs_fast_f32(lon.ctypes.data_as(fp), lat.ctypes.data_as(fp), satz.ctypes.data_as(fp),
                                          ctypes.c_long(ns), fine, out_lon.ctypes.data_as(fp),
                                          out_lat.ctypes.data_as(fp), modes)
        assert rc == 0
        return out_lon, out_lat, dict(zip(("slow", "low", "high", "both", "small", "eft_quads", "eft_flagged_rows",
                                           "eft_mismatches"), modes))
    run.f32 = run_f32

    def run_simple_f32(lon, lat, fine):
        ns, p = lon.shape[0] // 10, 1000 // fine
        out = [np.full((ns * 10 * p, 1354 * p), np.nan, np.float32) for _ in range(5)]
        modes = (ctypes.c_long * 6)()
        rc = lib.gtp_emul_simple_fast_f32(lon.ctypes.data_as(fp), lat.ctypes.data_as(fp), ctypes.c_long(ns), fine,
                                          *[o.ctypes.data_as(fp) for o in out], modes)
        assert rc == 0
        return out[0], out[1], tuple(out[2:]), dict(zip(("slow", "low", "high", "both", "small", "wide"), modes))
    run.simple_f32 = run_simple_f32

    def run_5km_f32(lon, lat, satz):
        ns, w = lon.shape[0] // 2, lon.shape[1]
        out_lon = np.full((ns * 10, 1354), np.nan, np.float32)
        out_lat = np.full_like(out_lon, np.nan)
        modes = (ctypes.c_long * 6)()
        rc = lib.gtp_emul_cviirs5_series_f32(lon.ctypes.data_as(fp), lat.ctypes.data_as(fp), satz.ctypes.data_as(fp),
                                             ctypes.c_long(ns), w, out_lon.ctypes.data_as(fp), out_lat.ctypes.data_as(fp), modes)
        assert rc == 0
        return out_lon, out_lat, dict(zip(("slow", "low", "high", "both", "small", "wide"), modes))
    run.five_km_f32 = run_5km_f32

    def run_xyz(lon, lat, satz, fine):
        ns, p = lon.shape[0] // 10, 1000 // fine
        out = [np.full((ns * 10 * p, 1354 * p), np.nan) for _ in range(3)]
        rc = lib.gtp_emul_xyz_fast_f64(lon.ctypes.data_as(dp), lat.ctypes.data_as(dp),
                                       satz.ctypes.data_as(dp) if satz is not None else None,
                                       ctypes.c_long(ns), fine, *[o.ctypes.data_as(dp) for o in out])
        assert rc == 0
        return tuple(out)
    run.xyz = run_xyz
    return run


@pytest.mark.parametrize("fine", [250, 500])
@pytest.mark.parametrize("granule", [0, 3, 5, 14, 287])
def test_fast_math_matches_oracle(emul, granule, fine):
    lon, lat, satz = testing.modis_granule(granule, n_scans=3)
    got_lon, got_lat, modes = emul(lon, lat, satz, fine)
    exp = gtp_oracle.cviirs(lon, lat, satz, 1000, fine, n_threads=4)
    assert_parity(got_lon, got_lat, *exp, what=f"host emulation granule {granule} -> {fine} m, modes {modes}")


def test_every_mode_is_exercised(emul):
    seen = {"slow": 0, "low": 0, "high": 0, "both": 0, "small": 0, "wide": 0}
    inputs = [testing.modis_granule(granule, n_scans=2) for granule in (0, 3, 5)]
    inputs.append(testing.stress_granule("pole", n_scans=2))      # the pole itself: rounding-faithful fallback
    for lon, lat, satz in inputs:
        for k, v in emul(lon, lat, satz, 250)[2].items():
            seen[k] += v
    assert all(v > 0 for v in seen.values()), seen


@pytest.mark.parametrize("name", ["antimeridian", "pole", "south_pole", "prime_meridian", "lat_switch",
                                  "symmetric_satz", "zeros", "nan_holes"])
def test_fast_math_stress_inputs(emul, golden_inputs, name):
    lon, lat, satz = (a.astype(np.float64) for a in golden_inputs[name])
    for fine in (250, 500):
        got_lon, got_lat, modes = emul(lon, lat, satz, fine)
        with np.errstate(all="ignore"):
            exp = gtp_oracle.cviirs(lon, lat, satz, 1000, fine)
        assert_parity(got_lon, got_lat, *exp, what=f"host emulation {name} -> {fine} m, modes {modes}")


@pytest.mark.parametrize("width", [271, 270])
@pytest.mark.parametrize("granule", [0, 3, 4, 5, 14, 34, 287])
def test_5km_fast_math_matches_oracle(emul, granule, width):
    """5 km -> 1 km: the 8-term series of csrc/gtp_fast5.cu against the oracle, whole granules (406 tie
    rows) incl. the ones that sweep over a pole."""
    lon, lat, satz = testing.subsample_5km(*testing.modis_granule(granule), width=width)
    got_lon, got_lat, modes = emul.five_km(lon, lat, satz)
    with np.errstate(all="ignore"):
        exp = gtp_oracle.cviirs(lon, lat, satz, 5000, 1000, width, n_threads=8)
    assert_parity(got_lon, got_lat, *exp, what=f"host emulation 5km({width}) granule {granule}, modes {modes}")


@pytest.mark.parametrize("name", ["antimeridian", "pole", "south_pole", "prime_meridian", "lat_switch",
                                  "symmetric_satz", "zeros", "nan_holes"])
def test_5km_fast_math_stress_inputs(emul, golden_inputs, name):
    from cases import cviirs_inputs
    for width in (271, 270):
        lon, lat, satz = cviirs_inputs(golden_inputs[name], 5000, width, np.float64)
        got_lon, got_lat, modes = emul.five_km(lon, lat, satz)
        with np.errstate(all="ignore"):
            exp = gtp_oracle.cviirs(lon, lat, satz, 5000, 1000, width)
        assert_parity(got_lon, got_lat, *exp, what=f"host emulation 5km({width}) {name}, modes {modes}")


def test_unphysical_inputs_fall_back(emul):
    """Tie points that are not MODIS-like spaced (random lon/lat) must take the slow path and
    still agree with the oracle."""
    rng = np.random.default_rng(5)
    lon = rng.uniform(-180, 180, (10, 1354))
    lat = rng.uniform(-89, 89, (10, 1354))
    satz = rng.uniform(0, 65, (10, 1354))
    got_lon, got_lat, modes = emul(lon, lat, satz, 250)
    assert modes["slow"] > 0.99 * sum(v for k, v in modes.items() if k != "small")
    with np.errstate(all="ignore"):
        exp = gtp_oracle.cviirs(lon, lat, satz, 1000, 250)
    # Chords between random points run deep inside the sphere, where the reference's formulas are
    # ill-conditioned in ways that never occur for real geolocation (rho -> 0 away from the poles,
    # rho / R -> 1 in the asin branch): an FMA contraction already moves the result by ~1e-9 deg.
    # This test is about the fallback being taken and being the right formula, so 1e-7 deg.
    from parity import compare
    s = compare(got_lon, got_lat, *exp)
    assert s["nan_pattern_equal"], s
    assert s["max_dlat"] < 1e-7 and max(s["max_dlon_outside_band"], s["max_dlon_in_band"]) < 1e-7, s


@pytest.mark.parametrize("fine", [250, 500])
@pytest.mark.parametrize("granule", [0, 3, 5, 14, 287])
def test_simple_fast_math_matches_oracle(emul, granule, fine):
    """simple_modis_interpolator as the bilinear form with c_exp = c_ali = 0 (SURVEY.md 8a
    closed form) against the oracle's map_coordinates + extrapolation restatement."""
    lon, lat, _ = testing.modis_granule(granule, n_scans=3)
    got_lon, got_lat, modes = emul.simple(lon, lat, fine)
    exp = gtp_oracle.simple(lon, lat, 1000, fine, n_threads=4)
    assert_parity(got_lon, got_lat, *exp, what=f"host emulation simple granule {granule} -> {fine} m, modes {modes}")


@pytest.mark.parametrize("name", ["antimeridian", "pole", "south_pole", "prime_meridian", "lat_switch", "zeros", "nan_holes"])
def test_simple_fast_math_stress_inputs(emul, golden_inputs, name):
    lon, lat, _ = (a.astype(np.float64) for a in golden_inputs[name])
    for fine in (250, 500):
        got_lon, got_lat, modes = emul.simple(lon, lat, fine)
        with np.errstate(all="ignore"):
            exp = gtp_oracle.simple(lon, lat, 1000, fine)
        assert_parity(got_lon, got_lat, *exp, what=f"host emulation simple {name} -> {fine} m, modes {modes}")


@pytest.mark.parametrize("fine", [250, 500])
@pytest.mark.parametrize("granule", [0, 3, 5, 14, 287])
def test_f32_fast_math_matches_oracle(emul, granule, fine):
    """float32 fast path: the blend reproduces the reference's rounding points, the back-transform
    is the anchor-relative series in double; the float32 result is (almost always) bit-identical."""
    from parity import compare
    lon, lat, satz = testing.modis_granule(granule, n_scans=3, dtype=np.float32)
    got_lon, got_lat, modes = emul.f32(lon, lat, satz, fine)
    exp = gtp_oracle.cviirs(lon, lat, satz, 1000, fine, n_threads=4)
    s = assert_parity(got_lon, got_lat, *exp, what=f"host emulation f32 granule {granule} -> {fine} m, modes {modes}")
    assert s["bit_exact_fraction"] > 0.999, s


@pytest.mark.parametrize("name", ["real", "antimeridian", "pole", "south_pole", "prime_meridian", "lat_switch",
                                  "symmetric_satz", "zeros", "nan_holes"])
def test_f32_fast_math_stress_inputs(emul, golden_inputs, name):
    lon, lat, satz = golden_inputs[name]
    for fine in (250, 500):
        got_lon, got_lat, modes = emul.f32(lon, lat, satz, fine)
        with np.errstate(all="ignore"):
            exp = gtp_oracle.cviirs(lon, lat, satz, 1000, fine)
        assert_parity(got_lon, got_lat, *exp, what=f"host emulation f32 {name} -> {fine} m, modes {modes}")


# ---- ECEF output (csrc/gtp_fast_xyz.cu) ----
def test_oracle_xyz_is_what_the_oracle_back_transforms():
    """The oracle's xyz output and its pinned lon / lat come from the same values: back-transforming the
    former with numpy's acos / asin (_modis_utils.pyx:43-65) reproduces the latter to libm rounding."""
    lon, lat, satz = testing.modis_granule(3, n_scans=2)
    x, y, z = gtp_oracle.cviirs_xyz(lon, lat, satz, 1000, 250)
    exp_lon, exp_lat = gtp_oracle.cviirs(lon, lat, satz, 1000, 250)
    rho = np.sqrt(x * x + y * y)
    got_lon = np.rad2deg(np.arccos(x / rho)) * np.sign(y)
    low = np.abs(z) < 0.8 * 6370997.0
    got_lat = np.where(low, 90.0 - np.rad2deg(np.arccos(z / 6370997.0)),
                       np.sign(z) * (90.0 - np.rad2deg(np.arcsin(np.minimum(rho / 6370997.0, 1.0)))))
    assert_parity(got_lon, got_lat, exp_lon, exp_lat, what="oracle xyz -> lon/lat")
    sx, sy, sz = gtp_oracle.simple_xyz(lon, lat, 1000, 250)
    assert np.abs(np.sqrt(sx * sx + sy * sy + sz * sz) - 6370997.0).max() < 5.0       # chords / extrapolation of <= 5 km quads


@pytest.mark.parametrize("fine", [250, 500])
@pytest.mark.parametrize("granule", [0, 3, 5, 287])
def test_xyz_fast_math_matches_oracle(emul, granule, fine):
    from parity import assert_xyz_parity
    lon, lat, satz = testing.modis_granule(granule, n_scans=3)
    assert_xyz_parity(emul.xyz(lon, lat, satz, fine), gtp_oracle.cviirs_xyz(lon, lat, satz, 1000, fine, n_threads=4),
                      what=f"host emulation xyz granule {granule} -> {fine} m")
    assert_xyz_parity(emul.xyz(lon, lat, None, fine), gtp_oracle.simple_xyz(lon, lat, 1000, fine, n_threads=4),
                      what=f"host emulation simple xyz granule {granule} -> {fine} m")


@pytest.mark.parametrize("name", ["antimeridian", "pole", "south_pole", "prime_meridian", "lat_switch",
                                  "symmetric_satz", "zeros", "nan_holes"])
def test_xyz_fast_math_stress_inputs(emul, golden_inputs, name):
    from parity import assert_xyz_parity
    lon, lat, satz = (a.astype(np.float64) for a in golden_inputs[name])
    for fine in (250, 500):
        with np.errstate(all="ignore"):
            exp = gtp_oracle.cviirs_xyz(lon, lat, satz, 1000, fine)
            exp_simple = gtp_oracle.simple_xyz(lon, lat, 1000, fine)
        assert_xyz_parity(emul.xyz(lon, lat, satz, fine), exp, what=f"host emulation xyz {name} -> {fine} m")
        assert_xyz_parity(emul.xyz(lon, lat, None, fine), exp_simple, what=f"host emulation simple xyz {name} -> {fine} m")


# ---- float32 simple interpolator fast path (simple_fast_1km_f32_kernel) ----
@pytest.mark.parametrize("fine", [250, 500])
@pytest.mark.parametrize("granule", [0, 3, 5, 14, 287])
def test_simple_f32_fast_math_matches_oracle(emul, granule, fine):
    """scipy's tap sum and the float32 column / row extrapolations are reproduced exactly: the float32
    fine xyz are bit-identical to the oracle's, lon / lat follow from the anchor-relative back-transform."""
    lon, lat, _ = testing.modis_granule(granule, n_scans=3, dtype=np.float32)
    got_lon, got_lat, xyz, modes = emul.simple_f32(lon, lat, fine)
    exp_xyz = gtp_oracle.simple_xyz(lon, lat, 1000, fine, n_threads=4)
    for g, e, name in zip(xyz, exp_xyz, "xyz"):
        assert np.array_equal(g, e), f"simple f32 {name} differs in {(g != e).sum()} pixels, max {np.abs(g - e).max()}"
    exp = gtp_oracle.simple(lon, lat, 1000, fine, n_threads=4)
    s = assert_parity(got_lon, got_lat, *exp, what=f"host emulation simple f32 granule {granule} -> {fine} m, modes {modes}")
    assert s["bit_exact_fraction"] > 0.999, s


@pytest.mark.parametrize("name", ["real", "antimeridian", "pole", "south_pole", "prime_meridian", "lat_switch", "zeros", "nan_holes"])
def test_simple_f32_fast_math_stress_inputs(emul, golden_inputs, name):
    lon, lat, _ = golden_inputs[name]
    for fine in (250, 500):
        got_lon, got_lat, xyz, modes = emul.simple_f32(lon, lat, fine)
        with np.errstate(all="ignore"):
            exp = gtp_oracle.simple(lon, lat, 1000, fine)
            exp_xyz = gtp_oracle.simple_xyz(lon, lat, 1000, fine)
        for g, e in zip(xyz, exp_xyz):
            assert np.array_equal(g, e, equal_nan=True), f"{name}: {(g != e).sum()} float32 xyz values differ"
        assert_parity(got_lon, got_lat, *exp, what=f"host emulation simple f32 {name} -> {fine} m, modes {modes}")


# ---- float32 5 km -> 1 km: series back-transform inside the strip kernel (gtp_generic.cu) ----
@pytest.mark.parametrize("width", [271, 270])
@pytest.mark.parametrize("granule", [0, 3, 4, 5, 14, 34, 287])
def test_5km_f32_series_matches_oracle(emul, granule, width):
    lon, lat, satz = testing.subsample_5km(*testing.modis_granule(granule, dtype=np.float32), width=width)
    got_lon, got_lat, modes = emul.five_km_f32(lon, lat, satz)
    with np.errstate(all="ignore"):
        exp = gtp_oracle.cviirs(lon, lat, satz, 5000, 1000, width, n_threads=8)
    s = assert_parity(got_lon, got_lat, *exp, what=f"host emulation 5km({width}) f32 granule {granule}, modes {modes}")
    assert s["bit_exact_fraction"] > 0.999, (s, modes)
    # modes["small"] counts pixels whose FP32-pipe blend (eft_pixel) differs from blend_f32 without having been flagged
    assert modes["small"] == 0, modes
    if granule in (0, 3, 287):          # away from the poles the series takes (nearly) every quad
        assert modes["slow"] < 0.02 * (modes["slow"] + modes["low"] + modes["high"] + modes["both"]), modes
        assert modes["wide"] > 0.8 * (modes["low"] + modes["high"] + modes["both"]), modes       # quads on the FP32 pipe


@pytest.mark.parametrize("name", ["real", "antimeridian", "pole", "south_pole", "prime_meridian", "lat_switch",
                                  "symmetric_satz", "zeros", "nan_holes"])
def test_5km_f32_series_stress_inputs(emul, golden_inputs, name):
    from cases import cviirs_inputs
    for width in (271, 270):
        lon, lat, satz = cviirs_inputs(golden_inputs[name], 5000, width, np.float32)
        got_lon, got_lat, modes = emul.five_km_f32(lon, lat, satz)
        with np.errstate(all="ignore"):
            exp = gtp_oracle.cviirs(lon, lat, satz, 5000, 1000, width)
        assert_parity(got_lon, got_lat, *exp, what=f"host emulation 5km({width}) f32 {name}, modes {modes}")


# ---- fill values in the geolocation ----
def assert_fill_parity(got, exp, rows_per_scan, what):
    """Scans 0 (valid) and 1 (all fill) to the usual tolerances; quads that MIX valid and fill corners are
    chords deep inside the sphere, where the reference's formulas are ill-conditioned (see
    test_unphysical_inputs_fall_back): 1e-7 deg there (float32: the usual 1e-5 deg / 1 ulp)."""
    from parity import compare
    n = 2 * rows_per_scan
    assert_parity(got[0][:n], got[1][:n], exp[0][:n], exp[1][:n], what=what + " (scans 0, 1)")
    s = compare(got[0], got[1], *exp)
    assert s["nan_pattern_equal"], (what, s)
    if got[0].dtype == np.float64:
        assert s["max_dlat"] < 1e-7 and max(s["max_dlon_outside_band"], s["max_dlon_in_band"]) < 1e-7, (what, s)
    else:
        assert s["n_over_tol_and_ulp"] == 0, (what, s)


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["float64", "float32"])
def test_fill_values_fold_like_the_reference(emul, dtype):
    """-999 (the MOD03 _FillValue) in lon / lat: the reference folds it through lonlat2xyz -> xyz2lonlat;
    the anchor-relative series must not keep it as the anchor (quads with an out-of-range anchor take the
    rounding-faithful path)."""
    lon, lat, satz = testing.fill_value_granule(dtype=dtype)
    name = np.dtype(dtype).name
    for fine in (250, 500):
        with np.errstate(all="ignore"):
            exp = gtp_oracle.cviirs(lon, lat, satz, 1000, fine)
            exp_s = gtp_oracle.simple(lon, lat, 1000, fine)
        assert np.nanmax(np.abs(exp[1])) <= 90.0 and np.nanmax(np.abs(exp[0])) <= 180.0      # the reference folds
        if dtype == np.float64:
            got_lon, got_lat, modes = emul(lon, lat, satz, fine)
            got_s = emul.simple(lon, lat, fine)
        else:
            got_lon, got_lat, modes = emul.f32(lon, lat, satz, fine)
            got_s = emul.simple_f32(lon, lat, fine)
        assert modes["slow"] > 0
        assert_fill_parity((got_lon, got_lat), exp, 10000 // fine, f"host emulation fill values -> {fine} m {name}")
        assert_fill_parity(got_s[:2], exp_s, 10000 // fine, f"host emulation simple fill values -> {fine} m {name}")
    for width in (271, 270):
        lon5, lat5, satz5 = testing.subsample_5km(lon, lat, satz, width)
        with np.errstate(all="ignore"):
            exp5 = gtp_oracle.cviirs(lon5, lat5, satz5, 5000, 1000, width)
        got5 = emul.five_km(lon5, lat5, satz5) if dtype == np.float64 else emul.five_km_f32(lon5, lat5, satz5)
        assert_fill_parity(got5[:2], exp5, 10, f"host emulation 5km({width}) fill values {name}")
