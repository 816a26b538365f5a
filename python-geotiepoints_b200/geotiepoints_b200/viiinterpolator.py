"""Interpolation of geographical tie points for the VII (EPS-SG METimage) products, on the GPU.

Drop-in for ``geotiepoints.viiinterpolator`` (``/root/reference/geotiepoints/viiinterpolator.py``; SURVEY.md 8f
row 4): ``tie_points_interpolation(data_on_tie_points, scan_alt_tie_points, tie_points_factor)`` (``:23-79``) and
``tie_points_geo_interpolation(longitude, latitude, scan_alt_tie_points, tie_points_factor,
lat_threshold_use_cartesian=60., z_threshold_use_xy=0.8)`` (``:83-134``) with the same arguments, results and
``ValueError``s.  Tie points are sub-sampled by ``tie_points_factor`` along and across the track; every scan of
``scan_alt_tie_points`` tie rows has its own edge tie points, so a pixel always blends two tie rows of its own
scan (``:55-62``).  The kernels are ``gtp_vii_interp_f64`` / ``gtp_vii_geo_interp_f64`` (``csrc/gtp_vii.cu``).

Containers: the reference takes ``xarray.DataArray``s; here numpy arrays (host path), float64 torch CUDA tensors
(device path, zero copy) and anything DataArray-like (``.dims`` + ``.values``: results are re-wrapped with the
same type and dims, coordinates dropped as the reference does, ``:75``).  Arithmetic and results are float64.
There is no CPU fallback.
"""
import ctypes

import numpy as np

from . import _capi
from ._operator import _is_torch_tensor


def _unwrap(arr):
    """(plain array, rewrap) for numpy / torch / DataArray-like input."""
    if hasattr(arr, "dims") and hasattr(arr, "values"):
        cls, dims = type(arr), tuple(arr.dims)
        return arr.values, (lambda out: cls(out, dims=dims))
    return arr, (lambda out: out)


def _shapes(shape, scan_alt_tie_points, tie_points_factor):
    n_tie_alt, n_tie_act = shape
    if n_tie_alt % scan_alt_tie_points != 0:
        # viiinterpolator.py:40-42
        raise ValueError("The number of tie points in the along-route dimension must be a multiple of %d"
                         % scan_alt_tie_points)
    n_scans = n_tie_alt // scan_alt_tie_points
    return (n_scans * (scan_alt_tie_points - 1) * tie_points_factor, (n_tie_act - 1) * tie_points_factor)


def tie_points_interpolation(data_on_tie_points, scan_alt_tie_points, tie_points_factor):
    """Interpolate a list of same-shaped arrays from the tie points to the pixel points (``:23-79``)."""
    lib = _capi.lib()
    S, F = int(scan_alt_tie_points), int(tie_points_factor)
    first, _ = _unwrap(data_on_tie_points[0])
    if len(first.shape) != 2:
        raise ValueError("Expected 2D tie-point arrays.")
    shape = tuple(first.shape)
    out_shape = _shapes(shape, S, F)
    dims0 = getattr(data_on_tie_points[0], "dims", None)
    results = []
    for item in data_on_tie_points:
        data, rewrap = _unwrap(item)
        if tuple(data.shape) != shape or getattr(item, "dims", None) != dims0:
            raise ValueError("The dimensions of the arrays are not consistent")          # :67-68
        if _is_torch_tensor(data):
            import torch
            if not data.is_cuda:
                raise TypeError("torch inputs must be CUDA tensors (numpy arrays take the host path)")
            data = data.to(torch.float64).contiguous()
            out = torch.empty(out_shape, dtype=torch.float64, device=data.device)
            stream = torch.cuda.current_stream(data.device).cuda_stream
            _capi.check(lib.gtp_vii_interp_f64(data.data_ptr(), shape[0], shape[1], S, F, out.data_ptr(),
                                               data.device.index, ctypes.c_void_p(stream)))
        else:
            data = np.ascontiguousarray(data, dtype=np.float64)
            out = _capi.pinned_pool.empty(out_shape, np.float64)
            _capi.check(lib.gtp_vii_interp_host_f64(data.ctypes.data, shape[0], shape[1], S, F, out.ctypes.data, -1))
        results.append(rewrap(out))
    return results


def tie_points_geo_interpolation(longitude, latitude, scan_alt_tie_points, tie_points_factor,
                                 lat_threshold_use_cartesian=60., z_threshold_use_xy=0.8):
    """Interpolate longitude / latitude from the tie points to the pixel points (``:83-134``)."""
    lib = _capi.lib()
    S, F = int(scan_alt_tie_points), int(tie_points_factor)
    lon, rewrap = _unwrap(longitude)
    lat, _ = _unwrap(latitude)
    if tuple(lon.shape) != tuple(lat.shape):
        raise ValueError("The dimensions of longitude and latitude don't match")          # :105-106
    shape = tuple(lon.shape)
    out_shape = _shapes(shape, S, F)
    if _is_torch_tensor(lon) and _is_torch_tensor(lat):
        import torch
        if not lon.is_cuda:
            raise TypeError("torch inputs must be CUDA tensors (numpy arrays take the host path)")
        lon, lat = lon.to(torch.float64).contiguous(), lat.to(torch.float64).contiguous()
        # :108-109 (two small reductions on the device, one scalar back)
        to_cart = bool(((lat.abs().max() > lat_threshold_use_cartesian) | ((lon.max() - lon.min()) > 180.)).item())
        out = [torch.empty(out_shape, dtype=torch.float64, device=lon.device) for _ in range(2)]
        stream = torch.cuda.current_stream(lon.device).cuda_stream
        _capi.check(lib.gtp_vii_geo_interp_f64(lon.data_ptr(), lat.data_ptr(), shape[0], shape[1], S, F, int(to_cart),
                                               float(z_threshold_use_xy), out[0].data_ptr(), out[1].data_ptr(),
                                               lon.device.index, ctypes.c_void_p(stream)))
        return out[0], out[1]
    lon = np.ascontiguousarray(lon, dtype=np.float64)
    lat = np.ascontiguousarray(lat, dtype=np.float64)
    to_cart = bool(np.max(np.fabs(lat)) > lat_threshold_use_cartesian or (np.max(lon) - np.min(lon)) > 180.)
    block = _capi.pinned_pool.empty((2,) + out_shape, np.float64)
    _capi.check(lib.gtp_vii_geo_interp_host_f64(lon.ctypes.data, lat.ctypes.data, shape[0], shape[1], S, F, int(to_cart),
                                                float(z_threshold_use_xy), block[0].ctypes.data, block[1].ctypes.data, -1))
    return rewrap(block[0]), rewrap(block[1])
