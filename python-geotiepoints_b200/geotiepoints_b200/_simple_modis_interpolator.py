"""Operator behind ``simple_modis_interpolator``: lon/lat-only bilinear interpolation.

Mirrors ``/root/reference/geotiepoints/_simple_modis_interpolator.pyx:13-64``: same
arguments and return value; the scan loop runs as one CUDA launch through
``gtp_simple_interp_*``.
"""
from . import _operator
from ._modis_utils import rows_per_scan_for_resolution


def interpolate_geolocation_cartesian(lon_array, lat_array, coarse_resolution, fine_resolution, ecef=False):
    arrays = [lon_array, lat_array]
    rows, cols = _operator._check_2d_same_shape(arrays)
    rows_per_scan = rows_per_scan_for_resolution(coarse_resolution)
    res_factor = coarse_resolution // fine_resolution
    if coarse_resolution != 1000 or res_factor not in (2, 4):
        raise ValueError("simple interpolation supports 1000 m -> 500 m or 250 m only")
    if rows % rows_per_scan:
        raise ValueError("Input longitude/latitude data does not consist of "
                         "whole scans (10 rows per scan).")
    n_scans = rows // rows_per_scan
    return _operator.run("simple", arrays, n_scans, (rows * res_factor, cols * res_factor), (cols, res_factor),
                         ecef=ecef)
