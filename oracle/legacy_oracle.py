"""CPU restatement of the reference's LEGACY spline interpolators (numpy only).

TEST INFRASTRUCTURE: only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` may
import it; the product package never does.

Restates ``modis5kmto1km`` / ``modis1kmto500m`` / ``modis1kmto250m``
(``/root/reference/geotiepoints/__init__.py:49-152``), i.e. ``GeoInterpolator``
(``geointerpolator.py:10-75``) over ``Interpolator`` (``interpolator.py:54-229``):

1. lon / lat -> Cartesian xyz on the sphere of radius 6370997 m (``geointerpolator.py:52-60``);
2. ``fill_borders("y", "x")`` (``interpolator.py:86-197``): per chunk of one scan, a first and a last row
   extrapolated linearly from the chunk's first / last two tie rows to the positions of the chunk's first /
   last fine row; then a first / last column extrapolated to the first / last fine column where the tie
   columns do not reach them;
3. ``RectBivariateSpline(rows, cols, data, s=0, kx=1, ky=3)`` (``:199-212``; kx = along-track order 1,
   ky = cross-track order 3): piecewise linear along the rows, and along the columns the interpolating
   cubic spline whose knots leave out the second and the second-to-last point - the not-a-knot spline;
4. xyz -> lon / lat (``geointerpolator.py:63-75``, thr = 0.8).

Because every step of 2 and the row direction of 3 are linear in the data with weights that do not depend on
the column, the border rows never have to be formed: a fine row between two tie rows of its scan is their
linear blend, a fine row above the first / below the last tie row of its scan is the linear extrapolation of
the scan's first / last two tie rows, and both commute with the column spline.  That is the form restated
here (and what the CUDA kernels do); it differs from the reference's order of operations by rounding only.
The not-a-knot spline is computed from its tridiagonal system for the second derivatives.

Pinned against the reference itself (``oracle/_ref/geotiepoints/legacy.py`` = the reference's module) in
``tests/test_legacy_oracle.py`` and against the reference's golden files
``testdata/test_5_to_1_geoloc_{5km,full}.h5`` (``geotiepoints/tests/test_modis.py:41-55``).
"""
import numpy as np

EARTH_RADIUS = 6370997.0

#: kind -> (tie rows per scan, fine rows per scan, tie-column positions, fine-column positions, row offset)
#: fine row j of a scan sits at tie-row coordinate (j + offset) / factor, tie row r at r (+ tie_row0)


def grids(kind):
    """(tie_cols, fine_cols, rows_per_scan, fine_rows_per_scan, tie_row_pos, fine_row_pos) of one scan."""
    if kind == "5km_1km":          # __init__.py:54-58
        return (np.arange(2, 1354, 5) / 5.0, np.arange(1354) / 5.0, 2, 10,
                np.arange(2, 10, 5) / 5.0, np.arange(10) / 5.0)
    if kind == "1km_500m":         # :106-110
        return (np.arange(1354.0), np.arange(1354 * 2) / 2.0, 10, 20,
                np.arange(10.0), (np.arange(20) - 0.5) / 2.0)
    if kind == "1km_250m":         # :133-141
        return (np.arange(1354.0), np.arange(1354 * 4) / 4.0, 10, 40,
                np.arange(10.0), (np.arange(40) - 1.5) / 4.0)
    raise ValueError(kind)


def lonlat2xyz(lons, lats):
    lons_rad, lats_rad = np.deg2rad(lons), np.deg2rad(lats)
    return (EARTH_RADIUS * np.cos(lats_rad) * np.cos(lons_rad),
            EARTH_RADIUS * np.cos(lats_rad) * np.sin(lons_rad),
            EARTH_RADIUS * np.sin(lats_rad))


def xyz2lonlat(x, y, z, thr=0.8):
    rho = np.sqrt(x ** 2 + y ** 2)
    lons = np.rad2deg(np.arccos(x / rho)) * np.sign(y)
    lats = np.sign(z) * (90 - np.rad2deg(np.arcsin(rho / EARTH_RADIUS)))
    nz = z / EARTH_RADIUS
    return lons, np.where(np.abs(nz) < thr, 90 - np.rad2deg(np.arccos(nz)), lats)


def extend_columns(data, tie_cols, fine_cols):
    """_fill_col_borders (:127-148): linear extrapolation to the first / last fine column where needed."""
    cols = tie_cols
    if tie_cols[0] != fine_cols[0]:
        first = data[:, 1] + (fine_cols[0] - tie_cols[1]) / (tie_cols[0] - tie_cols[1]) * (data[:, 0] - data[:, 1])
        data, cols = np.hstack((first[:, None], data)), np.concatenate(([fine_cols[0]], cols))
    if tie_cols[-1] != fine_cols[-1]:
        last = data[:, -1] + (fine_cols[-1] - tie_cols[-1]) / (tie_cols[-2] - tie_cols[-1]) * (data[:, -2] - data[:, -1])
        data, cols = np.hstack((data, last[:, None])), np.concatenate((cols, [fine_cols[-1]]))
    return data, cols


def notaknot_second_derivatives(x, y):
    """Second derivatives M (rows of y: one spline each) of the not-a-knot cubic spline through (x, y)."""
    n = x.size
    h = np.diff(x)
    d = np.diff(y, axis=1) / h
    rhs = 6.0 * (d[:, 1:] - d[:, :-1])                         # equations of the interior points 1 .. n-2
    lower, diag, upper = h[:-1].copy(), 2.0 * (h[:-1] + h[1:]), h[1:].copy()
    # not-a-knot: M0 = ((h0 + h1) M1 - h0 M2) / h1 and the mirror image at the other end, substituted
    diag[0] += h[0] * (h[0] + h[1]) / h[1]
    upper[0] -= h[0] * h[0] / h[1]
    diag[-1] += h[-1] * (h[-1] + h[-2]) / h[-2]
    lower[-1] -= h[-1] * h[-1] / h[-2]
    m = n - 2
    cp, dp = np.empty(m), np.empty((y.shape[0], m))
    cp[0], dp[:, 0] = upper[0] / diag[0], rhs[:, 0] / diag[0]
    for i in range(1, m):                                       # Thomas algorithm, all rows at once
        den = diag[i] - lower[i] * cp[i - 1]
        cp[i] = upper[i] / den
        dp[:, i] = (rhs[:, i] - lower[i] * dp[:, i - 1]) / den
    M = np.empty((y.shape[0], n))
    M[:, m] = dp[:, m - 1]
    for i in range(m - 2, -1, -1):
        M[:, i + 1] = dp[:, i] - cp[i] * M[:, i + 2]
    M[:, 0] = ((h[0] + h[1]) * M[:, 1] - h[0] * M[:, 2]) / h[1]
    M[:, -1] = ((h[-1] + h[-2]) * M[:, -2] - h[-1] * M[:, -3]) / h[-2]
    return M


def spline_columns(data, cols, fine_cols):
    """Not-a-knot cubic spline of every row of data at fine_cols."""
    M = notaknot_second_derivatives(cols, data)
    idx = np.clip(np.searchsorted(cols, fine_cols, side="right") - 1, 0, cols.size - 2)
    h = (cols[1:] - cols[:-1])[idx]
    t = fine_cols - cols[idx]
    y0, y1, m0, m1 = data[:, idx], data[:, idx + 1], M[:, idx], M[:, idx + 1]
    b = (y1 - y0) / h - h * (2.0 * m0 + m1) / 6.0
    return y0 + t * (b + t * (0.5 * m0 + t * (m1 - m0) / (6.0 * h)))


def row_weights(kind):
    """Per fine row of a scan: (upper tie row r, weight w) with value = (1 - w) tie[r] + w tie[r + 1]
    (w outside [0, 1] = the linear extrapolation of fill_borders('y'), interpolator.py:150-197)."""
    _, _, L, n, tie_pos, fine_pos = grids(kind)
    r = np.clip(np.searchsorted(tie_pos, fine_pos, side="right") - 1, 0, L - 2)
    w = (fine_pos - tie_pos[r]) / (tie_pos[r + 1] - tie_pos[r])
    return r, w


def interpolate_xyz(lon, lat, kind):
    tie_cols, fine_cols, L, n, _, _ = grids(kind)
    lon, lat = np.asarray(lon, np.float64), np.asarray(lat, np.float64)
    if lon.shape[0] % L or lon.shape[1] != tie_cols.size:
        raise ValueError(f"{kind}: expected whole scans of {L} rows x {tie_cols.size} columns, got {lon.shape}")
    scans = lon.shape[0] // L
    r, w = row_weights(kind)
    out = []
    for comp in lonlat2xyz(lon, lat):
        data, cols = extend_columns(comp, tie_cols, fine_cols)
        fine = spline_columns(data, cols, fine_cols).reshape(scans, L, fine_cols.size)
        blended = (1.0 - w)[None, :, None] * fine[:, r, :] + w[None, :, None] * fine[:, r + 1, :]
        out.append(blended.reshape(scans * n, fine_cols.size))
    return tuple(out)


def interpolate(lon, lat, kind):
    """(lons, lats) of the legacy interpolator `kind` ("5km_1km", "1km_500m", "1km_250m")."""
    with np.errstate(all="ignore"):
        return xyz2lonlat(*interpolate_xyz(lon, lat, kind))


def modis5kmto1km(lons5km, lats5km):
    return interpolate(lons5km, lats5km, "5km_1km")


def modis1kmto500m(lons1km, lats1km, cores=1):
    return interpolate(lons1km, lats1km, "1km_500m")


def modis1kmto250m(lons1km, lats1km, cores=1):
    return interpolate(lons1km, lats1km, "1km_250m")
