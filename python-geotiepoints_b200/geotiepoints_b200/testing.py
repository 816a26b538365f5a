"""Deterministic synthetic MODIS-shaped geolocation granules (numpy only).

Used by ``bench.py`` and ``tests/`` to make inputs of the shapes BASELINE.json names;
not part of the interpolation path.  The geometry follows SURVEY.md section 8(d):
spherical Earth R = 6370.997 km and a circular orbit at H = 709 km (the constants of
``/root/reference/geotiepoints/_modis_interpolator.pyx:10-12``), inclination 98.2
degrees, 1.4771 s per scan, 1354 frames over a +-55 degree scan, 10 detectors per scan
whose along-track footprint grows with slant range (the MODIS "bow-tie").
"""
import numpy as np

R_KM = 6370.997
H_KM = 709.0
INCLINATION_DEG = 98.2
SCAN_PERIOD_S = 1.4771
ORBIT_PERIOD_S = 98.8 * 60.0
FRAMES = 1354
DETECTORS = 10
SCANS_PER_GRANULE = 203
GRANULE_ARG_LAT_STEP_DEG = 18.22
GRANULE_NODE_SHIFT_DEG = -1.2534
EARTH_ROT_DEG_PER_S = 360.0 / 86164.1


def modis_granule(granule=0, n_scans=SCANS_PER_GRANULE, dtype=np.float64, jitter_deg=1e-4,
                  quantize_satz=False, node_lon_deg=0.0, arg_lat0_deg=None, inclination_deg=INCLINATION_DEG):
    """Return ``(lon, lat, satz)`` of shape ``(10 * n_scans, 1354)`` in degrees.

    ``granule`` selects the 5-minute slot of the day (0..287): the argument of latitude
    starts at ``granule * 18.22`` degrees and the ascending node moves west by 1.2534
    degrees per granule, so the 288 granules of a day cover every latitude, both polar
    caps and many antimeridian / prime-meridian crossings.  The random jitter (seeded by
    ``granule``) stands in for terrain correction and breaks exact satz symmetry.
    """
    rng = np.random.default_rng(int(granule))
    u0 = np.deg2rad(granule * GRANULE_ARG_LAT_STEP_DEG if arg_lat0_deg is None else arg_lat0_deg)
    node0 = np.deg2rad(node_lon_deg + granule * GRANULE_NODE_SHIFT_DEG)
    inc = np.deg2rad(inclination_deg)

    frame = np.arange(FRAMES, dtype=np.float64)
    theta = np.deg2rad((frame - 676.5) * (110.0 / FRAMES))           # scan angle at the satellite
    zen = np.arcsin((R_KM + H_KM) / R_KM * np.sin(theta))            # signed zenith angle at the pixel
    phi = zen - theta                                                # Earth-central angle across track
    slant = np.sqrt(R_KM ** 2 + (R_KM + H_KM) ** 2 - 2 * R_KM * (R_KM + H_KM) * np.cos(phi))
    growth = slant / H_KM                                            # bow-tie growth of the footprint

    scan = np.arange(n_scans, dtype=np.float64)
    det = np.arange(DETECTORS, dtype=np.float64)
    t = (scan[:, None] * SCAN_PERIOD_S + np.zeros(DETECTORS)[None, :]).reshape(-1)      # (rows,)
    u_centre = u0 + 2 * np.pi * t / ORBIT_PERIOD_S
    along_km = ((det - 4.5)[None, :] * np.ones(n_scans)[:, None]).reshape(-1)          # (rows,)
    u = u_centre[:, None] + (along_km[:, None] * growth[None, :]) / R_KM               # (rows, frames)

    # orbital frame: x to the ascending node, y along the motion at the node, z = orbit normal
    cphi, sphi = np.cos(phi)[None, :], np.sin(phi)[None, :]
    xo = cphi * np.cos(u)
    yo = cphi * np.sin(u)
    zo = sphi * np.ones_like(u)
    # tilt by the inclination about x, then rotate about z by the node longitude minus Earth rotation
    ye = yo * np.cos(inc) - zo * np.sin(inc)
    ze = yo * np.sin(inc) + zo * np.cos(inc)
    node = node0 - np.deg2rad(EARTH_ROT_DEG_PER_S) * t
    cn, sn = np.cos(node)[:, None], np.sin(node)[:, None]
    x = xo * cn - ye * sn
    y = xo * sn + ye * cn
    lon = np.rad2deg(np.arctan2(y, x))
    lat = np.rad2deg(np.arcsin(np.clip(ze, -1.0, 1.0)))
    satz = np.abs(np.rad2deg(zen))[None, :] * np.ones_like(lon)
    if jitter_deg:
        lon = lon + rng.normal(0.0, jitter_deg, lon.shape) / np.maximum(np.cos(np.deg2rad(lat)), 0.05)
        lat = np.clip(lat + rng.normal(0.0, jitter_deg, lat.shape), -90.0, 90.0)
        satz = np.abs(satz + rng.normal(0.0, jitter_deg, satz.shape))
    lon = (lon + 180.0) % 360.0 - 180.0
    if quantize_satz:
        satz = np.round(satz * 100.0) / 100.0     # MOD03 SensorZenith is int16 * 0.01
    return (np.ascontiguousarray(lon, dtype=dtype), np.ascontiguousarray(lat, dtype=dtype),
            np.ascontiguousarray(satz, dtype=dtype))


def subsample_5km(lon, lat, satz, width=271):
    """1 km arrays -> 5 km tie points the way the reference's tests do
    (``geotiepoints/tests/test_modisinterpolator.py:55-68``)."""
    sl = (slice(2, None, 5), slice(2, None, 5)) if width == 271 else (slice(2, None, 5), slice(2, -5, 5))
    return tuple(np.ascontiguousarray(a[sl]) for a in (lon, lat, satz))


def stress_granule(kind, n_scans=4, dtype=np.float64):
    """Small hand-placed granules for the edge cases of BASELINE config 5."""
    if kind == "antimeridian":
        lon, lat, satz = modis_granule(0, n_scans, np.float64, node_lon_deg=172.0, arg_lat0_deg=20.0)
    elif kind == "pole":
        # a polar orbit whose ground track crosses the pole inside the first scan
        lon, lat, satz = modis_granule(0, n_scans, np.float64, arg_lat0_deg=89.97, inclination_deg=90.0)
    elif kind == "south_pole":
        lon, lat, satz = modis_granule(0, n_scans, np.float64, arg_lat0_deg=269.97, inclination_deg=90.0)
    elif kind == "prime_meridian":
        lon, lat, satz = modis_granule(0, n_scans, np.float64, node_lon_deg=-1.0, arg_lat0_deg=1.0)
    elif kind == "lat_switch":           # straddles |lat| = 53.13 deg (z = 0.8 R), _modis_utils.pyx:62
        lon, lat, satz = modis_granule(0, n_scans, np.float64, arg_lat0_deg=54.4)
    elif kind == "symmetric_satz":       # GH #19, tests/test_modisinterpolator.py:143-149
        lon, lat, _ = modis_granule(0, n_scans, np.float64)
        satz = np.abs(np.linspace(-65.4, 65.4, FRAMES, dtype=np.float32)).astype(np.float64)
        satz = np.tile(satz, (lon.shape[0], 1))
    elif kind == "zeros":                # BASELINE config 1 (all-zero lon/lat fixture)
        lon = np.zeros((10 * n_scans, FRAMES)); lat = np.zeros_like(lon)
        _, _, satz = modis_granule(0, n_scans, np.float64)
    else:
        raise ValueError(kind)
    return tuple(np.ascontiguousarray(a, dtype=dtype) for a in (lon, lat, satz))


STRESS_KINDS = ("antimeridian", "pole", "south_pole", "prime_meridian", "lat_switch", "symmetric_satz", "zeros")


def punch_nan_holes(lon, lat, satz, seed=7, n=24):
    """Copy of the inputs with a few NaN tie points (lon/lat pairs and satz separately);
    the reference propagates them only to the pixels of the touching quads."""
    rng = np.random.default_rng(seed)
    lon, lat, satz = lon.copy(), lat.copy(), satz.copy()
    rows, cols = lon.shape
    for k in range(n):
        r, c = int(rng.integers(0, rows)), int(rng.integers(0, cols))
        if k % 2:
            lon[r, c] = np.nan
            lat[r, c] = np.nan
        else:
            satz[r, c] = np.nan
    # always include the corners and edges
    lon[0, 0] = lat[0, 0] = np.nan
    satz[rows - 1, cols - 1] = np.nan
    return lon, lat, satz


def fill_value_granule(granule=3, n_scans=4, dtype=np.float64, fill=-999.0):
    """A granule with the MOD03 ``_FillValue`` (-999) in its geolocation: scan 1 entirely, a block of
    columns of scan 2 and a few single tie points of scan 3.  The reference has no notion of fill values:
    it folds them through ``lonlat2xyz`` -> ``xyz2lonlat`` like any other angle (-999 deg comes back as
    81 deg), and so must every path here."""
    lon, lat, satz = modis_granule(granule, n_scans=n_scans, dtype=dtype)
    lon, lat = lon.copy(), lat.copy()
    lon[10:20], lat[10:20] = fill, fill
    if n_scans > 2:
        lon[20:30, 400:460], lat[20:30, 400:460] = fill, fill
    if n_scans > 3:
        for r, c in ((31, 5), (34, 700), (39, 1353), (30, 1352), (35, 0)):
            lon[r, c], lat[r, c] = fill, fill
    return lon, lat, satz
