"""Interpolation of MODIS geolocation using the satellite zenith angle (CVIIRS scheme).

Public API identical to ``geotiepoints.modisinterpolator``
(``/root/reference/geotiepoints/modisinterpolator.py:21-62``).
"""
import warnings

from ._modis_interpolator import interpolate


def modis_1km_to_250m(lon1, lat1, satz1):
    """Interpolate MODIS geolocation from 1km to 250m resolution."""
    return interpolate(lon1, lat1, satz1, coarse_resolution=1000, fine_resolution=250)


def modis_1km_to_500m(lon1, lat1, satz1):
    """Interpolate MODIS geolocation from 1km to 500m resolution."""
    return interpolate(lon1, lat1, satz1, coarse_resolution=1000, fine_resolution=500)


def modis_5km_to_1km(lon1, lat1, satz1):
    """Interpolate MODIS geolocation from 5km to 1km resolution."""
    return interpolate(lon1, lat1, satz1, coarse_resolution=5000, fine_resolution=1000,
                       coarse_scan_width=lon1.shape[1])


def modis_5km_to_500m(lon1, lat1, satz1):
    """Interpolate MODIS geolocation from 5km to 500m resolution."""
    warnings.warn("Interpolating 5km geolocation to 500m resolution may result in poor quality")
    return interpolate(lon1, lat1, satz1, coarse_resolution=5000, fine_resolution=500,
                       coarse_scan_width=lon1.shape[1])


def modis_5km_to_250m(lon1, lat1, satz1):
    """Interpolate MODIS geolocation from 5km to 250m resolution."""
    warnings.warn("Interpolating 5km geolocation to 250m resolution may result in poor quality")
    return interpolate(lon1, lat1, satz1, coarse_resolution=5000, fine_resolution=250,
                       coarse_scan_width=lon1.shape[1])
