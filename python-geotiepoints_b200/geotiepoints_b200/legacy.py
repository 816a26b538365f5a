"""The legacy spline interpolators of ``geotiepoints`` (package level functions) on the GPU.

Drop-in for ``geotiepoints.modis5kmto1km`` / ``modis1kmto500m`` / ``modis1kmto250m`` and
``get_scene_splits`` (``/root/reference/geotiepoints/__init__.py:15-152``): same names,
arguments and return values.  The reference runs ``GeoInterpolator`` (``geointerpolator.py:10-75``) over
``Interpolator`` (``interpolator.py:54-229``) with along-track order 1, cross-track order 3,
``chunk_size`` = one scan and ``fill_borders("y", "x")``, i.e. scipy's
``RectBivariateSpline(kx=1, ky=3, s=0)`` on the Cartesian tie points; here that is
``gtp_legacy_interp_f64`` (``csrc/gtp_legacy.cu``): a not-a-knot cubic spline along the tie columns (one
tridiagonal solve per tie row and Cartesian component), linear blending / extrapolation along the rows of a
scan, and the reference's back-transform.

Inputs: 2-D arrays (numpy, or float64 torch CUDA tensors for the zero-copy device path) of whole scans -
``(2 n, 271)`` for 5 km, ``(10 n, 1354)`` for 1 km.  Like scipy's spline, the arithmetic is float64 whatever
the input dtype (integer and float32 arrays are widened; the reference's float32 rounding of the Cartesian
tie points for float32 input is not reproduced) and the results are float64.  ``cores`` is accepted and
ignored: the reference uses it to split the swath over ``multiprocessing`` workers
(``__init__.py:74-94``), the GPU does not need it.  There is no CPU fallback.
"""
import ctypes

import numpy as np

from . import _capi
from ._operator import _is_torch_tensor, empty_outputs, empty_outputs_torch

_KINDS = {"5km_1km": (0, 2, 271, 5, 1354), "1km_500m": (1, 10, 1354, 2, 2708), "1km_250m": (2, 10, 1354, 4, 5416)}


def get_scene_splits(nlines_swath, nlines_scan, n_cpus):
    """Line numbers at which the reference splits a swath for its worker pool (``__init__.py:15-28``)."""
    nscans = nlines_swath // nlines_scan
    nscans_subscene = 1 if nscans < n_cpus else nscans // n_cpus
    nlines_subscene = nscans_subscene * nlines_scan
    return range(nlines_subscene, nlines_swath, nlines_subscene)


def modis5kmto1km(lons5km, lats5km):
    """1 km geolocation for MODIS from 5 km tie points (``__init__.py:49-71``)."""
    return _interpolate(lons5km, lats5km, "5km_1km")


def modis1kmto500m(lons1km, lats1km, cores=1):
    """500 m geolocation for MODIS from 1 km tie points (``__init__.py:97-122``)."""
    return _interpolate(lons1km, lats1km, "1km_500m")


def modis1kmto250m(lons1km, lats1km, cores=1):
    """250 m geolocation for MODIS from 1 km tie points (``__init__.py:125-152``)."""
    return _interpolate(lons1km, lats1km, "1km_250m")


def _interpolate(lons, lats, kind):
    code, rows_per_scan, cols, factor, out_cols = _KINDS[kind]
    lib = _capi.lib()
    if _is_torch_tensor(lons) and _is_torch_tensor(lats):
        import torch
        if not lons.is_cuda:
            raise TypeError("torch inputs must be CUDA tensors (numpy arrays take the host path)")
        lons, lats = lons.to(torch.float64).contiguous(), lats.to(torch.float64).contiguous()
        shape = _check(tuple(lons.shape), tuple(lats.shape), rows_per_scan, cols, kind)
        out = empty_outputs_torch(2, (shape[0] * factor, out_cols), torch.float64, lons.device)
        stream = torch.cuda.current_stream(lons.device).cuda_stream
        _capi.check(lib.gtp_legacy_interp_f64(lons.data_ptr(), lats.data_ptr(), shape[0] // rows_per_scan, code,
                                              out[0].data_ptr(), out[1].data_ptr(), lons.device.index,
                                              ctypes.c_void_p(stream)))
        return out
    lons = np.ascontiguousarray(lons, dtype=np.float64)
    lats = np.ascontiguousarray(lats, dtype=np.float64)
    shape = _check(lons.shape, lats.shape, rows_per_scan, cols, kind)
    out = empty_outputs(2, (shape[0] * factor, out_cols), np.float64)
    _capi.check(lib.gtp_legacy_interp_host_f64(lons.ctypes.data, lats.ctypes.data, shape[0] // rows_per_scan, code,
                                               out[0].ctypes.data, out[1].ctypes.data, -1))
    return out


def _check(shape, other, rows_per_scan, cols, kind):
    if len(shape) != 2 or shape != other:
        raise ValueError(f"expected two 2-D arrays of the same shape, got {shape} and {other}")
    if shape[0] % rows_per_scan or shape[1] != cols:
        # the reference fails later and less clearly here (index errors in fill_borders / the spline)
        raise ValueError(f"{kind}: expected whole scans of {rows_per_scan} rows x {cols} columns, got {shape}")
    return shape
