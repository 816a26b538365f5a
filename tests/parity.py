"""Parity metrics shared by the tests, smoke() and bench.py.

Tolerances (BASELINE.json north_star): <= 1e-10 deg for float64, <= 1e-5 deg for float32.
Two honest caveats from SURVEY.md section 0.4, both measured on the reference itself:

* float64 longitude is ill-conditioned in two places, and the reference's own output is
  noise there.  (a) Where |sin(lon)| is tiny: the reference evaluates ``acos(r)``,
  ``r = x / sqrt(x^2 + y^2)`` (``_modis_utils.pyx:58``).  A perturbation of r by K ulps
  (u = 2^-53; the rounding of x, y, the sqrt and the division) moves the result by
  K u / |sin(lon)| radians, and by up to sqrt(2 K u) radians where r rounds to +-1 (the
  reference cannot represent 0 < |lon| < 8.5e-7 deg at all).  (b) Next to the poles: x and y
  carry at least one ulp(R) = 9.3e-10 m of rounding noise, which is K ulp(R) / rho radians
  of longitude at distance rho from the axis - more than 1e-10 deg within ~4 km of a pole.
  Outside these two regions the 1e-10 deg tolerance is enforced as is; inside them the bound
  is ``conditioning_tolerance`` with K = 16 (a) / K = 8 (b) and the statistics are reported
  separately ("band").
* one float32 ulp at |lon| >= 128 deg is 1.53e-5 deg > 1e-5 deg, so a single last-bit
  flip of a correctly computed value already exceeds the stated tolerance.  float32
  results are accepted when they are within 1e-5 deg OR within one float32 ulp.
"""
import numpy as np

TOL_F64 = 1e-10
TOL_F32 = 1e-5
BAND = 2e-4          # |sin(lon)| below this = conditioning band (a)
BAND_K_ULPS = 16.0
POLE_K_ULPS = 8.0
_U = 2.0 ** -53
_ULP_R = float(np.spacing(6370997.0))
_R = 6370997.0


def conditioning_tolerance(lon_deg, lat_deg):
    """Largest |dlon| (deg) explained by rounding noise in the reference's own x, y, r."""
    s = np.abs(np.sin(np.deg2rad(lon_deg.astype(np.float64))))
    rho = _R * np.abs(np.cos(np.deg2rad(lat_deg.astype(np.float64))))
    with np.errstate(divide="ignore"):
        meridian = np.minimum(BAND_K_ULPS * _U / s, np.sqrt(2.0 * BAND_K_ULPS * _U))
        pole = np.minimum(POLE_K_ULPS * _ULP_R / rho, np.pi)
    return np.maximum(np.rad2deg(np.maximum(meridian, pole)), TOL_F64)


def lon_diff(a, b):
    """|a - b| on the circle, degrees."""
    d = np.abs(a.astype(np.float64) - b.astype(np.float64))
    return np.minimum(d, 360.0 - d)


def compare(got_lon, got_lat, exp_lon, exp_lat):
    """Return a dict of parity statistics between a result and the oracle's."""
    assert got_lon.shape == exp_lon.shape and got_lat.shape == exp_lat.shape
    assert got_lon.dtype == exp_lon.dtype
    nan_same = bool(np.array_equal(np.isnan(got_lon), np.isnan(exp_lon))
                    and np.array_equal(np.isnan(got_lat), np.isnan(exp_lat)))
    ok = ~(np.isnan(exp_lon) | np.isnan(exp_lat) | np.isnan(got_lon) | np.isnan(got_lat))
    dlon = np.where(ok, lon_diff(got_lon, exp_lon), 0.0)
    dlat = np.where(ok, np.abs(got_lat.astype(np.float64) - exp_lat.astype(np.float64)), 0.0)
    band = np.abs(np.sin(np.deg2rad(exp_lon.astype(np.float64)))) < BAND
    band |= np.rad2deg(POLE_K_ULPS * _ULP_R / np.maximum(_R * np.abs(np.cos(np.deg2rad(exp_lat.astype(np.float64)))), 1e-30)) > TOL_F64
    band &= ok
    stats = {
        "n": int(ok.sum()),
        "nan_pattern_equal": nan_same,
        "max_dlat": float(dlat.max(initial=0.0)),
        "max_dlon_outside_band": float(np.where(band, 0.0, dlon).max(initial=0.0)),
        "max_dlon_in_band": float(np.where(band, dlon, 0.0).max(initial=0.0)),
        "n_in_band_over_conditioning_bound": int((band & (dlon > conditioning_tolerance(exp_lon, exp_lat))).sum()),
        "n_in_band": int(band.sum()),
        "bit_exact_fraction": float(((got_lon == exp_lon) & (got_lat == exp_lat))[ok].mean()) if ok.any() else 1.0,
    }
    if got_lon.dtype == np.float32:
        ulp_lon = np.spacing(np.abs(exp_lon).astype(np.float32)).astype(np.float64)
        ulp_lat = np.spacing(np.abs(exp_lat).astype(np.float32)).astype(np.float64)
        bad = ((dlon > TOL_F32) & (dlon > ulp_lon * 1.0001)) | ((dlat > TOL_F32) & (dlat > ulp_lat * 1.0001))
        stats["n_over_tol_and_ulp"] = int((bad & ok).sum())
        stats["n_over_tol"] = int((((dlon > TOL_F32) | (dlat > TOL_F32)) & ok).sum())
        stats["max_dlon"] = float(dlon.max(initial=0.0))
    return stats


def assert_parity(got_lon, got_lat, exp_lon, exp_lat, what=""):
    s = compare(got_lon, got_lat, exp_lon, exp_lat)
    assert s["nan_pattern_equal"], f"{what}: NaN pattern differs from the oracle: {s}"
    if got_lon.dtype == np.float64:
        assert s["max_dlat"] <= TOL_F64, f"{what}: lat off by {s['max_dlat']:.3e} deg > {TOL_F64}: {s}"
        assert s["max_dlon_outside_band"] <= TOL_F64, f"{what}: lon off by {s['max_dlon_outside_band']:.3e} deg: {s}"
        assert s["n_in_band_over_conditioning_bound"] == 0, \
            f"{what}: lon inside a conditioning band beyond the rounding noise of the reference itself: {s}"
    else:
        assert s["n_over_tol_and_ulp"] == 0, f"{what}: {s['n_over_tol_and_ulp']} float32 pixels beyond 1e-5 deg and 1 ulp: {s}"
    return s


def haversine_m(lon1, lat1, lon2, lat2, radius=6371008.8):
    """Great-circle distance in metres.  Stands in for ``pyproj.Geod(ellps='WGS84').inv``
    of the reference's tests (tests/test_modisinterpolator.py:85-107); pyproj is not
    installed.  The two differ by < 0.6 %, well inside the slack checked here."""
    lon1, lat1, lon2, lat2 = (np.deg2rad(np.asarray(a, dtype=np.float64)) for a in (lon1, lat1, lon2, lat2))
    a = np.sin((lat2 - lat1) / 2) ** 2 + np.cos(lat1) * np.cos(lat2) * np.sin((lon2 - lon1) / 2) ** 2
    return 2 * radius * np.arcsin(np.sqrt(np.clip(a, 0.0, 1.0)))


# ---- ECEF output (metres) ----
#: float64: the north_star tolerance of 1e-10 deg is 1.1e-5 m of arc on the reference's sphere; the CUDA
#: path is held to 1e-5 m.  Measured over the whole synthetic day (profiles/r01_full_day_parity.txt):
#: <= 1.5e-6 m for CVIIRS, <= 1.2e-8 m for the simple interpolator.  The CVIIRS figure is the reference's
#: own conditioning: its c_expansion divides two differences of nearly equal angles (:62-64), so next
#: to nadir a_scan carries up to ~1e-9 of noise, times a quad width of 1-5 km.  The worst pixel of the
#: day (granule 72, nadir quad, zenith angles 1.1e-7 deg apart) moves by 5e-7 m in the ORACLE when the
#: zenith angles are perturbed by one ulp; typical quads agree to 4e-8 m, and the rounding-faithful
#: kernel (CUDA libm vs glibc) is as far from the oracle as the fast one.
#: float32: one float32 ulp of R is 0.5 m; the rounding-faithful kernels reproduce the reference's
#: rounding points, so at most 1 ulp (a double rounding of a libm result that differs in its last bit).
TOL_XYZ_F64_M = 1e-5


def compare_xyz(got, exp):
    """Statistics of an (x, y, z) triple against the oracle's, metres."""
    n_nan_diff, max_abs, max_ulps, exact, n = 0, 0.0, 0.0, 0, 0
    for g, e in zip(got, exp):
        assert g.shape == e.shape and g.dtype == e.dtype
        n_nan_diff += int((np.isnan(g) != np.isnan(e)).sum())
        ok = ~(np.isnan(g) | np.isnan(e))
        d = np.abs(g[ok].astype(np.float64) - e[ok].astype(np.float64))
        max_abs = max(max_abs, float(d.max(initial=0.0)))
        ulp = np.spacing(np.maximum(np.abs(e[ok]), np.asarray(1.0, e.dtype))).astype(np.float64)
        max_ulps = max(max_ulps, float((d / ulp).max(initial=0.0)))
        exact += int((g[ok] == e[ok]).sum())
        n += int(ok.sum())
    return {"n": n, "nan_mismatches": n_nan_diff, "max_abs_m": max_abs, "max_ulps": max_ulps,
            "bit_exact_fraction": exact / n if n else 1.0}


def assert_xyz_parity(got, exp, what=""):
    s = compare_xyz(got, exp)
    assert s["nan_mismatches"] == 0, (what, s)
    if got[0].dtype == np.float64:
        assert s["max_abs_m"] <= TOL_XYZ_F64_M, (what, s)
    else:
        assert s["max_ulps"] <= 1.0, (what, s)
    return s
