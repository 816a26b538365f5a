"""ctypes loader for the CPU oracle (``oracle/libgtp_oracle.so``).  TEST INFRASTRUCTURE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this module; the product package never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgtp_oracle.so")
_lib = None


def build(force=False):
    src = [os.path.join(_HERE, n) for n in ("gtp_oracle.c", "gtp_oracle_body.inc")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libgtp_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        for suf, ct in (("f32", ctypes.c_float), ("f64", ctypes.c_double)):
            p = ctypes.POINTER(ct)
            fn = getattr(_lib, f"gtp_oracle_cviirs_{suf}")
            fn.restype = ctypes.c_int
            fn.argtypes = [p, p, p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int, p, p, ctypes.c_int]
            fn = getattr(_lib, f"gtp_oracle_simple_{suf}")
            fn.restype = ctypes.c_int
            fn.argtypes = [p, p, ctypes.c_long, ctypes.c_long, ctypes.c_int, p, p, ctypes.c_int]
            fn = getattr(_lib, f"gtp_oracle_cviirs_xyz_{suf}")
            fn.restype = ctypes.c_int
            fn.argtypes = [p, p, p, ctypes.c_long, ctypes.c_int, ctypes.c_int, ctypes.c_int, p, p, p, ctypes.c_int]
            fn = getattr(_lib, f"gtp_oracle_simple_xyz_{suf}")
            fn.restype = ctypes.c_int
            fn.argtypes = [p, p, ctypes.c_long, ctypes.c_long, ctypes.c_int, p, p, p, ctypes.c_int]
    return _lib


def _ptr(a):
    ct = ctypes.c_float if a.dtype == np.float32 else ctypes.c_double
    return a.ctypes.data_as(ctypes.POINTER(ct))


def _suffix(dtype):
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    raise TypeError(f"oracle supports float32/float64, got {dtype}")


def cviirs(lon, lat, satz, coarse_resolution, fine_resolution, coarse_scan_width=0, n_threads=1):
    """Oracle of ``_modis_interpolator.interpolate`` (``_modis_interpolator.pyx:17-38``)."""
    lon, lat, satz = (np.ascontiguousarray(a) for a in (lon, lat, satz))
    assert lon.dtype == lat.dtype == satz.dtype and lon.shape == lat.shape == satz.shape
    L = 10 if coarse_resolution == 1000 else 2
    f = (1000 // fine_resolution) * (coarse_resolution // 1000)
    n_scans = lon.shape[0] // L
    Wf = 1354 * (1000 // fine_resolution)
    if coarse_resolution == 5000:
        coarse_scan_width = coarse_scan_width or lon.shape[1]
    out_lon = np.empty((lon.shape[0] * f, Wf), dtype=lon.dtype)
    out_lat = np.empty_like(out_lon)
    fn = getattr(lib(), "gtp_oracle_cviirs_" + _suffix(lon.dtype))
    rc = fn(_ptr(lon), _ptr(lat), _ptr(satz), n_scans, coarse_resolution, fine_resolution,
            coarse_scan_width, _ptr(out_lon), _ptr(out_lat), n_threads)
    if rc:
        raise ValueError(f"oracle cviirs rc={rc}")
    return out_lon, out_lat


def simple(lon, lat, coarse_resolution, fine_resolution, n_threads=1):
    """Oracle of ``_simple_modis_interpolator.interpolate_geolocation_cartesian``
    (``_simple_modis_interpolator.pyx:13-64``)."""
    lon, lat = (np.ascontiguousarray(a) for a in (lon, lat))
    assert lon.dtype == lat.dtype and lon.shape == lat.shape
    res = coarse_resolution // fine_resolution
    n_scans = lon.shape[0] // 10
    out_lon = np.empty((lon.shape[0] * res, lon.shape[1] * res), dtype=lon.dtype)
    out_lat = np.empty_like(out_lon)
    fn = getattr(lib(), "gtp_oracle_simple_" + _suffix(lon.dtype))
    rc = fn(_ptr(lon), _ptr(lat), n_scans, lon.shape[1], res, _ptr(out_lon), _ptr(out_lat), n_threads)
    if rc:
        raise ValueError(f"oracle simple rc={rc}")
    return out_lon, out_lat


def cviirs_xyz(lon, lat, satz, coarse_resolution, fine_resolution, coarse_scan_width=0, n_threads=1):
    """Fine ECEF coordinates (metres) of ``_compute_fine_xyz`` (``_modis_interpolator.pyx:346-398``)."""
    lon, lat, satz = (np.ascontiguousarray(a) for a in (lon, lat, satz))
    assert lon.dtype == lat.dtype == satz.dtype and lon.shape == lat.shape == satz.shape
    L = 10 if coarse_resolution == 1000 else 2
    f = (1000 // fine_resolution) * (coarse_resolution // 1000)
    if coarse_resolution == 5000:
        coarse_scan_width = coarse_scan_width or lon.shape[1]
    out = [np.empty((lon.shape[0] * f, 1354 * (1000 // fine_resolution)), dtype=lon.dtype) for _ in range(3)]
    fn = getattr(lib(), "gtp_oracle_cviirs_xyz_" + _suffix(lon.dtype))
    rc = fn(_ptr(lon), _ptr(lat), _ptr(satz), lon.shape[0] // L, coarse_resolution, fine_resolution,
            coarse_scan_width, *[_ptr(o) for o in out], n_threads)
    if rc:
        raise ValueError(f"oracle cviirs_xyz rc={rc}")
    return tuple(out)


def simple_xyz(lon, lat, coarse_resolution, fine_resolution, n_threads=1):
    """Fine ECEF coordinates of the simple interpolator (``_simple_modis_interpolator.pyx:52-63``)."""
    lon, lat = (np.ascontiguousarray(a) for a in (lon, lat))
    assert lon.dtype == lat.dtype and lon.shape == lat.shape
    res = coarse_resolution // fine_resolution
    out = [np.empty((lon.shape[0] * res, lon.shape[1] * res), dtype=lon.dtype) for _ in range(3)]
    fn = getattr(lib(), "gtp_oracle_simple_xyz_" + _suffix(lon.dtype))
    rc = fn(_ptr(lon), _ptr(lat), lon.shape[0] // 10, lon.shape[1], res, *[_ptr(o) for o in out], n_threads)
    if rc:
        raise ValueError(f"oracle simple_xyz rc={rc}")
    return tuple(out)
