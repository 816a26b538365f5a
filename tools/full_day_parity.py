#!/usr/bin/env python
"""Parity over the whole synthetic day: 3 scans out of every one of the 288 granules (864 scans, 1.9e8
pixels at 250 m) through the CUDA path and through the oracle, per configuration.

    python tools/full_day_parity.py            # needs a GPU; prints one line per configuration
    python tools/full_day_parity.py ecef       # only the ECEF-output configurations

The oracle (oracle/, test infrastructure) is the checker here, exactly as in tests/."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in ("python-geotiepoints_b200", "oracle", "tests"):
    sys.path.insert(0, os.path.join(ROOT, p))
import numpy as np
import gtp_oracle
from geotiepoints_b200 import ecef, mod03, modisinterpolator, simple_modis_interpolator, testing
from parity import compare, compare_xyz

only_ecef = len(sys.argv) > 1 and sys.argv[1] == "ecef"

scans = (0, 101, 202)
parts = []
t0 = time.time()
for g in range(288):
    lon, lat, satz = testing.modis_granule(g)
    rows = np.concatenate([np.arange(s * 10, s * 10 + 10) for s in scans])
    parts.append((lon[rows], lat[rows], satz[rows]))
lon, lat, satz = (np.ascontiguousarray(np.concatenate([p[k] for p in parts])) for k in range(3))
print(f"# {lon.shape[0] // 10} scans of 288 granules, |lat| max {np.abs(lat).max():.2f}, built in {time.time() - t0:.0f} s", flush=True)


def report(name, got, exp):
    s = compare(got[0], got[1], exp[0], exp[1])
    keys = ("n", "nan_pattern_equal", "max_dlat", "max_dlon_outside_band", "max_dlon_in_band",
            "n_in_band_over_conditioning_bound", "bit_exact_fraction", "n_over_tol_and_ulp")
    print(f"{name:34s} " + " ".join(f"{k}={s[k]:.3g}" if isinstance(s[k], float) else f"{k}={s[k]}" for k in keys if k in s), flush=True)


def report_xyz(name, got, exp):
    s = compare_xyz(got, exp)
    print(f"{name:34s} " + " ".join(f"{k}={v:.3g}" if isinstance(v, float) else f"{k}={v}" for k, v in s.items()), flush=True)


with np.errstate(all="ignore"):
    # ECEF output (metres): float64 fast kernel, float32 rounding-faithful kernel, MOD03 types
    for fine, fn, sfn in ((250, ecef.modis_1km_to_250m, ecef.simple_modis_1km_to_250m),
                          (500, ecef.modis_1km_to_500m, ecef.simple_modis_1km_to_500m)):
        for dt in (np.float64, np.float32):
            a = [x.astype(dt) for x in (lon, lat, satz)]
            report_xyz(f"ecef cviirs 1km->{fine}m {np.dtype(dt).name}", fn(*a), gtp_oracle.cviirs_xyz(*a, 1000, fine, 0, n_threads=16))
        a = [lon, lat]
        report_xyz(f"ecef simple 1km->{fine}m float64", sfn(*a), gtp_oracle.simple_xyz(*a, 1000, fine, n_threads=16))
    f = [x.astype(np.float32) for x in (lon, lat)]
    zen = np.round(satz / 0.01).astype(np.int16)
    exp = gtp_oracle.cviirs_xyz(f[0].astype(np.float64), f[1].astype(np.float64), zen.astype(np.float64) * 0.01, 1000, 250, 0, n_threads=16)
    report_xyz("ecef mod03 f32+i16 -> f64", mod03.modis_1km_to_250m(*f, zen, zenith_scale=0.01, out_dtype=np.float64, ecef=True), exp)
    report_xyz("ecef mod03 f32+i16 -> f32", mod03.modis_1km_to_250m(*f, zen, zenith_scale=0.01, ecef=True),
               tuple(x.astype(np.float32) for x in exp))
    if only_ecef:
        sys.exit(0)
    for fine, fn in ((250, modisinterpolator.modis_1km_to_250m), (500, modisinterpolator.modis_1km_to_500m)):
        for dt in (np.float64, np.float32):
            a = [x.astype(dt) for x in (lon, lat, satz)]
            report(f"cviirs 1km->{fine}m {np.dtype(dt).name}", fn(*a), gtp_oracle.cviirs(*a, 1000, fine, 0, n_threads=16))
    for fine, fn in ((250, simple_modis_interpolator.modis_1km_to_250m), (500, simple_modis_interpolator.modis_1km_to_500m)):
        for dt in (np.float64, np.float32):
            a = [x.astype(dt) for x in (lon, lat)]
            report(f"simple 1km->{fine}m {np.dtype(dt).name}", fn(*a), gtp_oracle.simple(*a, 1000, fine, n_threads=16))
    # 5 km: every scan of 12 granules spread over the day (a 5 km scan is 2 tie rows of its own)
    for width in (271, 270):
        for dt in (np.float64, np.float32):
            subs = [testing.subsample_5km(*[x.astype(dt) for x in testing.modis_granule(g)], width=width) for g in range(0, 288, 24)]
            a = [np.ascontiguousarray(np.concatenate([s[k] for s in subs])) for k in range(3)]
            report(f"cviirs 5km({width})->1km {np.dtype(dt).name}", modisinterpolator.modis_5km_to_1km(*a),
                   gtp_oracle.cviirs(*a, 5000, 1000, width, n_threads=16))
    # extension: float32 lon/lat + int16 zenith, float64 math -> float32
    f = [x.astype(np.float32) for x in (lon, lat)]
    zen = np.round(satz / 0.01).astype(np.int16)
    exp = gtp_oracle.cviirs(f[0].astype(np.float64), f[1].astype(np.float64), zen.astype(np.float64) * 0.01, 1000, 250, 0, n_threads=16)
    report("mod03 1km->250m f32+i16 -> f64", mod03.modis_1km_to_250m(*f, zen, zenith_scale=0.01, out_dtype=np.float64), exp)
    report("mod03 1km->250m f32+i16 -> f32", mod03.modis_1km_to_250m(*f, zen, zenith_scale=0.01),
           tuple(x.astype(np.float32) for x in exp))
    # the legacy spline interpolators against the numpy restatement (oracle/legacy_oracle.py)
    import legacy_oracle
    from geotiepoints_b200 import legacy
    report("legacy 1km->250m float64", legacy.modis1kmto250m(lon, lat), legacy_oracle.interpolate(lon, lat, "1km_250m"))
    report("legacy 1km->500m float64", legacy.modis1kmto500m(lon, lat), legacy_oracle.interpolate(lon, lat, "1km_500m"))
    subs = [testing.modis_granule(g) for g in range(0, 288, 24)]
    l5 = [np.ascontiguousarray(np.concatenate([s[k][2::5, 2::5] for s in subs])) for k in range(2)]
    report("legacy 5km->1km float64", legacy.modis5kmto1km(*l5), legacy_oracle.interpolate(*l5, "5km_1km"))
    # ... the VII interpolator (every scan has its own tie rows: the 3 sampled scans of a granule are 3 VII scans of 10
    # tie rows; Cartesian mode) and the generic grid classes (the sampled scans of one granule are one rectilinear grid
    # each: tie rows 4 r + 1.5, columns 4 c, extrapolating to the 40 x 5416 pixels of the scan)
    import grid_oracle
    import vii_oracle
    from geotiepoints_b200 import viiinterpolator
    from geotiepoints_b200.geointerpolator import GeoGridInterpolator, GeoSplineInterpolator
    for factor in (4, 8, 3):
        got = viiinterpolator.tie_points_geo_interpolation(lon, lat, 10, factor, lat_threshold_use_cartesian=-1.0)
        xyz = vii_oracle.tie_points_interpolation(list(vii_oracle.lonlat2xyz(lon, lat)), 10, factor)
        report(f"vii geo (Cartesian) factor {factor}", got, vii_oracle.xyz2lonlat(*xyz, 0.8))
    ty, tx, fy, fx = 4.0 * np.arange(10) + 1.5, 4.0 * np.arange(1354), np.arange(40.0), np.arange(5416.0)
    for name, klass, kwargs, okw in (("linear", GeoGridInterpolator, dict(bounds_error=False, fill_value=None), dict(method_y=1, method_x=1)),
                                     ("cubic", GeoGridInterpolator, dict(bounds_error=False, fill_value=None, method="cubic"), dict(method_y=3, method_x=3)),
                                     ("spline kx=1 ky=3", GeoSplineInterpolator, dict(kx=1, ky=3), dict(method_y=1, method_x=3, oob=grid_oracle.CLAMP))):
        got, exp = [[], []], [[], []]
        for s in range(0, lon.shape[0] // 10, 3):              # one sampled scan per granule
            rows = slice(10 * s, 10 * s + 10)
            g_ = klass((ty, tx), lon[rows], lat[rows], **kwargs).interpolate((fy, fx))
            e_ = grid_oracle.geogrid_interpolate(ty, tx, lon[rows], lat[rows], fy, fx, **okw)
            for k in (0, 1):
                got[k].append(g_[k]); exp[k].append(e_[k])
        report(f"geogrid {name}", [np.concatenate(x) for x in got], [np.concatenate(x) for x in exp])
    # exact-equality counts of the float32 paths (the verdict's "array_equal" question): pixels that differ at all
    for fine, fn in ((250, modisinterpolator.modis_1km_to_250m), (500, modisinterpolator.modis_1km_to_500m)):
        a = [x.astype(np.float32) for x in (lon, lat, satz)]
        got, exp = fn(*a), gtp_oracle.cviirs(*a, 1000, fine, 0, n_threads=16)
        nd = int(((got[0] != exp[0]) | (got[1] != exp[1])).sum())
        d = np.maximum(np.abs(got[0].astype(np.float64) - exp[0]), np.abs(got[1].astype(np.float64) - exp[1]))
        print(f"cviirs 1km->{fine}m float32: {nd} of {got[0].size} pixels differ from the oracle at all "
              f"({nd / got[0].size:.2e}); of those beyond 1e-6 deg: {int((d > 1e-6).sum())}, max {d.max():.3g} deg", flush=True)
