"""Interpolate MODIS 1km navigation arrays to 250m and 500m resolutions (lon/lat only).

Public API identical to ``geotiepoints.simple_modis_interpolator``
(``/root/reference/geotiepoints/simple_modis_interpolator.py:21-61``).
"""
from ._modis_utils import scanline_mapblocks
from ._simple_modis_interpolator import interpolate_geolocation_cartesian as _interp_cuda


@scanline_mapblocks
def interpolate_geolocation_cartesian(lon_array, lat_array, coarse_resolution, fine_resolution):
    """Interpolate MODIS navigation from 1000m resolution to 250m or 500m.

    Arguments:
        lon_array: Longitude data as a 2D numpy, dask, or xarray DataArray object
            (or a torch CUDA tensor) representing 1000m geolocation.
        lat_array: Latitude data, same container and shape.

    Returns:
        A two-element tuple (lon, lat).
    """
    return _interp_cuda(lon_array, lat_array, coarse_resolution, fine_resolution)


def modis_1km_to_250m(lon1, lat1):
    """Interpolate MODIS geolocation from 1km to 250m resolution."""
    return interpolate_geolocation_cartesian(lon1, lat1, coarse_resolution=1000, fine_resolution=250)


def modis_1km_to_500m(lon1, lat1):
    """Interpolate MODIS geolocation from 1km to 500m resolution."""
    return interpolate_geolocation_cartesian(lon1, lat1, coarse_resolution=1000, fine_resolution=500)
