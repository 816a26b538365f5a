"""CPU restatement of the reference's VII (EPS-SG METimage) tie-point interpolation (numpy only).

TEST INFRASTRUCTURE: only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` may import it.

Restates ``/root/reference/geotiepoints/viiinterpolator.py``:
* ``tie_points_interpolation`` (``:23-79``): per scan of ``scan_alt_tie_points`` tie rows the pixel grid is
  ``(S - 1) F`` rows x ``(n_act - 1) F`` columns; two linear ``xarray.interp`` calls, first along then across the
  track (``:72-75``), i.e. scipy ``interp1d``'s ``y_lo + (y_hi - y_lo) / (x_hi - x_lo) * (x - x_lo)`` twice.  The
  pixel rows never fall between two scans (``:55-62``), so a pixel always blends two tie rows of its own scan.
* ``tie_points_geo_interpolation`` (``:83-134``): on Cartesian coordinates (sphere of the IUGG mean radius,
  ``:20``) when any |latitude| > 60 deg or the longitudes span more than 180 deg, else on lon / lat directly;
  ``_lonlat2xyz`` (``:137-154``), ``_xyz2lonlat`` (``:157-178``, z threshold 0.8).

PARITY: xarray is not installed in this image, so the reference's functions themselves cannot be run here
("unpinned" against an execution of the reference).  The restatement is pinned against the golden vectors the
reference's own tests hold (``geotiepoints/tests/test_viiinterpolator.py:22-96``, 8 printed digits,
``np.allclose``), re-typed in ``tests/test_vii_oracle.py``.
"""
import numpy as np

MEAN_EARTH_RADIUS = 6371008.7714   # viiinterpolator.py:20


def tie_points_interpolation(data_on_tie_points, scan_alt_tie_points, tie_points_factor):
    first = np.asarray(data_on_tie_points[0])
    n_tie_alt, n_tie_act = first.shape
    S, F = int(scan_alt_tie_points), int(tie_points_factor)
    if n_tie_alt % S != 0:
        raise ValueError("The number of tie points in the along-route dimension must be a multiple of %d" % S)
    n_scans = n_tie_alt // S
    rows_local = np.arange((S - 1) * F)
    r0 = (np.arange(n_scans)[:, None] * S + rows_local[None, :] // F).ravel()          # upper tie row of every pixel row
    wr = np.tile((rows_local % F) / float(F), n_scans)
    cols = np.arange((n_tie_act - 1) * F)
    c0, wc = cols // F, (cols % F) / float(F)
    out = []
    for data in data_on_tie_points:
        data = np.asarray(data, dtype=np.float64)
        if data.shape != (n_tie_alt, n_tie_act):
            raise ValueError("The dimensions of the arrays are not consistent")
        along = data[r0] + (data[r0 + 1] - data[r0]) * wr[:, None]                     # first along the track ...
        out.append(along[:, c0] + (along[:, c0 + 1] - along[:, c0]) * wc[None, :])     # ... then across
    return out


def lonlat2xyz(lons, lats):
    lons_rad, lats_rad = np.deg2rad(lons), np.deg2rad(lats)
    return (MEAN_EARTH_RADIUS * np.cos(lats_rad) * np.cos(lons_rad),
            MEAN_EARTH_RADIUS * np.cos(lats_rad) * np.sin(lons_rad),
            MEAN_EARTH_RADIUS * np.sin(lats_rad))


def xyz2lonlat(x, y, z, z_threshold_use_xy=0.8):
    r = np.sqrt(x ** 2 + y ** 2)
    thr_z = z_threshold_use_xy * MEAN_EARTH_RADIUS
    lons = np.rad2deg(np.arccos(x / r)) * np.sign(y)
    lats = np.where(np.logical_and(z < thr_z, z > -thr_z),
                    90. - np.rad2deg(np.arccos(z / MEAN_EARTH_RADIUS)),
                    np.sign(z) * (90. - np.rad2deg(np.arcsin(r / MEAN_EARTH_RADIUS))))
    return lons, lats


def use_cartesian(longitude, latitude, lat_threshold_use_cartesian=60.):
    return bool(np.max(np.fabs(latitude)) > lat_threshold_use_cartesian or (np.max(longitude) - np.min(longitude)) > 180.)


def tie_points_geo_interpolation(longitude, latitude, scan_alt_tie_points, tie_points_factor,
                                 lat_threshold_use_cartesian=60., z_threshold_use_xy=0.8):
    longitude, latitude = np.asarray(longitude, np.float64), np.asarray(latitude, np.float64)
    if longitude.shape != latitude.shape:
        raise ValueError("The dimensions of longitude and latitude don't match")
    if use_cartesian(longitude, latitude, lat_threshold_use_cartesian):
        xyz = tie_points_interpolation(list(lonlat2xyz(longitude, latitude)), scan_alt_tie_points, tie_points_factor)
        with np.errstate(all="ignore"):
            return xyz2lonlat(*xyz, z_threshold_use_xy)
    lon, lat = tie_points_interpolation([longitude, latitude], scan_alt_tie_points, tie_points_factor)
    return lon, lat
