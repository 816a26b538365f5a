"""MODIS geolocation interpolation straight from the MOD03 on-disk types (extension).

Not part of the reference's API.  MOD03 stores ``Longitude`` /
``Latitude`` as float32 and ``SensorZenith`` as int16 with ``scale_factor = 0.01``
(``/root/reference/testdata/create_modis_test_data.py:85-99``); a caller of the reference widens
them on the host and gets the reference's float32 arithmetic, whose fine Cartesian coordinates are
float32 (0.25 - 0.5 m of rounding noise, up to 4e-5 deg of longitude at high latitudes).  The
functions here hand the arrays to the GPU as they are, decode the zenith angle in the kernel, run
the **float64** algorithm (the kernel behind ``modisinterpolator.modis_1km_to_250m`` for float64
input) and round once, to float32 by default.  The result is the reference's *float64* result for
``lon.astype(float64), lat.astype(float64), zenith * scale + offset`` - less input and half the
output traffic of the float64 path, three times the pixel rate of the float32-faithful one.

    lons, lats = mod03.modis_1km_to_250m(lon_f32, lat_f32, sensor_zenith_i16, zenith_scale=0.01)

Inputs are numpy arrays or torch CUDA tensors (zero copy), C-contiguous, whole scans of 10 rows and
1354 columns; ``sensor_zenith`` is float32 (degrees, or scaled) or int16.  ``zenith_scale`` defaults to the
MOD03 ``scale_factor`` 0.01 for int16 and to 1 for float32; fill values (-999, -32767) are taken as they are,
like any other number (the reference has no notion of them either).

``ecef=True`` (1 km tie points) returns the Cartesian ``(x, y, z)`` in metres that the algorithm
interpolates - see :mod:`geotiepoints_b200.ecef` - instead of their back-transform to lon / lat.
"""
import ctypes

import numpy as np

from . import _capi
from ._modis_utils import scanline_mapblocks
from ._operator import _is_torch_tensor, empty_outputs, empty_outputs_torch

_ZEN_KIND = {"float32": 0, "int16": 1}
_OUT_KIND = {"float32": 0, "float64": 1}


def modis_1km_to_250m(lon1, lat1, sensor_zenith, *, zenith_scale=None, zenith_offset=0.0, out_dtype=np.float32, ecef=False):
    """1 km -> 250 m from float32 lon / lat and float32 | int16 zenith angles, float64 math."""
    return _dispatch(lon1, lat1, sensor_zenith, 1000, 250, zenith_scale, zenith_offset, out_dtype, ecef)


def modis_1km_to_500m(lon1, lat1, sensor_zenith, *, zenith_scale=None, zenith_offset=0.0, out_dtype=np.float32, ecef=False):
    """1 km -> 500 m from float32 lon / lat and float32 | int16 zenith angles, float64 math."""
    return _dispatch(lon1, lat1, sensor_zenith, 1000, 500, zenith_scale, zenith_offset, out_dtype, ecef)


def modis_5km_to_1km(lon1, lat1, sensor_zenith, *, zenith_scale=None, zenith_offset=0.0, out_dtype=np.float32):
    """5 km (270 or 271 columns) -> 1 km from float32 lon / lat and float32 | int16 zenith angles, float64 math."""
    return _dispatch(lon1, lat1, sensor_zenith, 5000, 1000, zenith_scale, zenith_offset, out_dtype, False)


def simple_modis_1km_to_250m(lon1, lat1, *, out_dtype=np.float32, ecef=False):
    """``simple_modis_interpolator.modis_1km_to_250m`` on float32 lon / lat with float64 math."""
    return _dispatch_simple(lon1, lat1, 250, out_dtype, ecef)


def simple_modis_1km_to_500m(lon1, lat1, *, out_dtype=np.float32, ecef=False):
    """``simple_modis_interpolator.modis_1km_to_500m`` on float32 lon / lat with float64 math."""
    return _dispatch_simple(lon1, lat1, 500, out_dtype, ecef)


# dask / xarray inputs: lazily, whole-scan block by block, through the same adapter as the drop-in operators
@scanline_mapblocks
def _lonlat_blocks(lon1, lat1, zen, coarse_resolution=0, fine_resolution=0, zenith_scale=None, zenith_offset=0.0,
                   out_dtype=np.float32):
    return _interpolate(lon1, lat1, zen, fine_resolution, zenith_scale, zenith_offset, out_dtype,
                        coarse_resolution=coarse_resolution)


@scanline_mapblocks(n_outputs=3)
def _xyz_blocks(lon1, lat1, zen, coarse_resolution=0, fine_resolution=0, zenith_scale=None, zenith_offset=0.0,
                out_dtype=np.float32):
    return _interpolate(lon1, lat1, zen, fine_resolution, zenith_scale, zenith_offset, out_dtype,
                        coarse_resolution=coarse_resolution, ecef=True)


@scanline_mapblocks
def _simple_lonlat_blocks(lon1, lat1, coarse_resolution=0, fine_resolution=0, out_dtype=np.float32):
    return _interpolate_simple(lon1, lat1, coarse_resolution // fine_resolution, out_dtype, False)


@scanline_mapblocks(n_outputs=3)
def _simple_xyz_blocks(lon1, lat1, coarse_resolution=0, fine_resolution=0, out_dtype=np.float32):
    return _interpolate_simple(lon1, lat1, coarse_resolution // fine_resolution, out_dtype, True)


def _dispatch(lon1, lat1, zen, coarse, fine, zenith_scale, zenith_offset, out_dtype, ecef):
    fn = _xyz_blocks if ecef else _lonlat_blocks
    return fn(lon1, lat1, zen, coarse_resolution=coarse, fine_resolution=fine, zenith_scale=zenith_scale,
              zenith_offset=zenith_offset, out_dtype=out_dtype)


def _dispatch_simple(lon1, lat1, fine, out_dtype, ecef):
    fn = _simple_xyz_blocks if ecef else _simple_lonlat_blocks
    return fn(lon1, lat1, coarse_resolution=1000, fine_resolution=fine, out_dtype=out_dtype)


def _dtype_name(a):
    return str(a.dtype).replace("torch.", "")


def _interpolate(lon1, lat1, zen, fine_resolution, zenith_scale, zenith_offset, out_dtype, coarse_resolution=1000,
                 ecef=False):
    arrays = [lon1, lat1, zen]
    for a in arrays:
        if a.ndim != 2:
            raise ValueError("Expected 2D input arrays.")
    if _dtype_name(lon1) != "float32" or _dtype_name(lat1) != "float32":
        raise TypeError("mod03: longitude and latitude must be float32 (use modisinterpolator for float64)")
    if _dtype_name(zen) not in _ZEN_KIND:
        raise TypeError("mod03: sensor_zenith must be float32 or int16")
    if zenith_scale is None:
        # MOD03 SensorZenith: int16 with scale_factor 0.01; float32 zenith angles are degrees already
        zenith_scale = 0.01 if _dtype_name(zen) == "int16" else 1.0
    out_name = np.dtype(out_dtype).name
    if out_name not in _OUT_KIND:
        raise TypeError("mod03: out_dtype must be float32 or float64")
    shape = tuple(lon1.shape)
    if tuple(lat1.shape) != shape or tuple(zen.shape) != shape:
        raise ValueError("input arrays must have the same shape")
    rows, cols = shape
    rows_per_scan = 10 if coarse_resolution == 1000 else 2
    if rows % rows_per_scan:
        raise ValueError(f"Input data does not consist of whole scans ({rows_per_scan} rows per scan), got {rows} rows.")
    if coarse_resolution == 1000:
        if cols != 1354:
            raise ValueError(f"Expected 1354 tie-point columns, got {cols}.")
        p = 1000 // fine_resolution
        out_shape = (rows * p, cols * p)
    else:
        if cols not in (270, 271):
            raise NotImplementedError("Can't interpolate if 5km tiepoints have less than 270 columns.")
        out_shape = (rows * 5, 1354)
    scalars = (_ZEN_KIND[_dtype_name(zen)], float(zenith_scale), float(zenith_offset), rows // rows_per_scan,
               coarse_resolution, fine_resolution)
    if not ecef:                # the ECEF operator is 1 km only and has no coarse_width argument
        scalars += (cols if coarse_resolution == 5000 else 0,)
    n_out = 3 if ecef else 2
    name = "gtp_cviirs_interp_xyz_mixed" if ecef else "gtp_cviirs_interp_mixed"
    lib = _capi.lib()
    if all(_is_torch_tensor(a) for a in arrays):
        import torch
        if not lon1.is_cuda:
            raise TypeError("torch inputs must be CUDA tensors (numpy arrays take the host path)")
        arrays = [a.contiguous() for a in arrays]
        tdtype = torch.float32 if out_name == "float32" else torch.float64
        out = empty_outputs_torch(n_out, out_shape, tdtype, lon1.device)
        stream = torch.cuda.current_stream(lon1.device).cuda_stream
        rc = getattr(lib, name)(*[a.data_ptr() for a in arrays], *scalars, *[o.data_ptr() for o in out],
                                _OUT_KIND[out_name], lon1.device.index, ctypes.c_void_p(stream))
        _capi.check(rc)
        return tuple(out)
    arrays = [np.ascontiguousarray(a) for a in arrays]
    out = empty_outputs(n_out, out_shape, np.dtype(out_dtype))
    rc = getattr(lib, name + "_host")(*[a.ctypes.data for a in arrays], *scalars, *[o.ctypes.data for o in out],
                                      _OUT_KIND[out_name], -1)
    _capi.check(rc)
    return tuple(out)


def _interpolate_simple(lon1, lat1, res, out_dtype, ecef=False):
    arrays = [lon1, lat1]
    for a in arrays:
        if a.ndim != 2:
            raise ValueError("Expected 2D input arrays.")
    if _dtype_name(lon1) != "float32" or _dtype_name(lat1) != "float32":
        raise TypeError("mod03: longitude and latitude must be float32 (use simple_modis_interpolator for float64)")
    out_name = np.dtype(out_dtype).name
    if out_name not in _OUT_KIND:
        raise TypeError("mod03: out_dtype must be float32 or float64")
    rows, cols = tuple(lon1.shape)
    if tuple(lat1.shape) != (rows, cols):
        raise ValueError("input arrays must have the same shape")
    if rows % 10:
        raise ValueError(f"Input data does not consist of whole scans (10 rows per scan), got {rows} rows.")
    if cols != 1354:
        raise ValueError(f"Expected 1354 tie-point columns, got {cols}.")
    out_shape = (rows * res, cols * res)
    n_out = 3 if ecef else 2
    name = "gtp_simple_interp_xyz_mixed" if ecef else "gtp_simple_interp_mixed"
    lib = _capi.lib()
    if all(_is_torch_tensor(a) for a in arrays):
        import torch
        if not lon1.is_cuda:
            raise TypeError("torch inputs must be CUDA tensors (numpy arrays take the host path)")
        arrays = [a.contiguous() for a in arrays]
        tdtype = torch.float32 if out_name == "float32" else torch.float64
        out = empty_outputs_torch(n_out, out_shape, tdtype, lon1.device)
        stream = torch.cuda.current_stream(lon1.device).cuda_stream
        rc = getattr(lib, name)(*[a.data_ptr() for a in arrays], rows // 10, res, *[o.data_ptr() for o in out],
                                _OUT_KIND[out_name], lon1.device.index, ctypes.c_void_p(stream))
        _capi.check(rc)
        return tuple(out)
    arrays = [np.ascontiguousarray(a) for a in arrays]
    out = empty_outputs(n_out, out_shape, np.dtype(out_dtype))
    rc = getattr(lib, name + "_host")(*[a.ctypes.data for a in arrays], rows // 10, res, *[o.ctypes.data for o in out],
                                      _OUT_KIND[out_name], -1)
    _capi.check(rc)
    return tuple(out)
