"""ECEF output of the MODIS interpolators (extension).

The reference interpolates in Cartesian coordinates and converts the result back to longitude /
latitude as its last step (``xyz2lonlat``, ``/root/reference/geotiepoints/_modis_interpolator.pyx:293``,
``_modis_utils.pyx:43-65``); what usually follows - resampling with a kd-tree or a gradient search -
converts them straight back.  The functions here return the fine Cartesian coordinates themselves,
``(x, y, z)`` in metres on the reference's sphere of radius 6370997 m (``_modis_utils.pyx:23``): the
values of ``MODISInterpolator._compute_fine_xyz`` (``_modis_interpolator.pyx:346-398``) and of
``new_xyz`` in ``_simple_modis_interpolator.pyx:52-63``, in the arithmetic of the input dtype.

    x, y, z = ecef.modis_1km_to_250m(lon1, lat1, satz1)

Same argument rules as :mod:`geotiepoints_b200.modisinterpolator`: 2-D float32 or float64 arrays of one
dtype, whole scans, numpy arrays (host path) or torch CUDA tensors (zero copy); dask arrays (bare or
inside xarray ``DataArray``s) are mapped lazily, whole-scan block by block, by the same
``scanline_mapblocks`` adapter as the lon / lat operators.  The MOD03 on-disk types are served by
``mod03.modis_1km_to_250m(..., ecef=True)``.
"""
import warnings

from ._modis_interpolator import MODISInterpolator
from ._modis_utils import scanline_mapblocks
from ._simple_modis_interpolator import interpolate_geolocation_cartesian as _simple_operator


@scanline_mapblocks(n_outputs=3)
def _interpolate_blocks(lon1, lat1, satz1, coarse_resolution=0, fine_resolution=0, coarse_scan_width=0):
    return MODISInterpolator(coarse_resolution, fine_resolution, coarse_scan_width or 0).interpolate_xyz(lon1, lat1, satz1)


@scanline_mapblocks(n_outputs=3)
def _simple_blocks(lon1, lat1, coarse_resolution=0, fine_resolution=0):
    return _simple_operator(lon1, lat1, coarse_resolution, fine_resolution, ecef=True)


def _simple(lon1, lat1, coarse_resolution, fine_resolution, ecef=True):
    return _simple_blocks(lon1, lat1, coarse_resolution=coarse_resolution, fine_resolution=fine_resolution)


def interpolate(lon1, lat1, satz1, coarse_resolution, fine_resolution, coarse_scan_width=0):
    """ECEF counterpart of ``_modis_interpolator.interpolate`` (``:17-38``)."""
    if coarse_resolution == 5000 and coarse_scan_width not in (0, 270, 271):
        raise NotImplementedError("Can't interpolate if 5km tiepoints have less than 270 columns.")
    return _interpolate_blocks(lon1, lat1, satz1, coarse_resolution=coarse_resolution, fine_resolution=fine_resolution,
                               coarse_scan_width=coarse_scan_width)


def modis_1km_to_250m(lon1, lat1, satz1):
    """1 km -> 250 m, ``(x, y, z)`` in metres."""
    return interpolate(lon1, lat1, satz1, 1000, 250)


def modis_1km_to_500m(lon1, lat1, satz1):
    """1 km -> 500 m, ``(x, y, z)`` in metres."""
    return interpolate(lon1, lat1, satz1, 1000, 500)


def modis_5km_to_1km(lon1, lat1, satz1):
    """5 km (270 | 271 columns) -> 1 km, ``(x, y, z)`` in metres."""
    return interpolate(lon1, lat1, satz1, 5000, 1000, lon1.shape[1])


def modis_5km_to_500m(lon1, lat1, satz1):
    """5 km -> 500 m, ``(x, y, z)`` in metres."""
    warnings.warn("Interpolating 5km geolocation to 500m resolution may result in poor quality")
    return interpolate(lon1, lat1, satz1, 5000, 500, lon1.shape[1])


def modis_5km_to_250m(lon1, lat1, satz1):
    """5 km -> 250 m, ``(x, y, z)`` in metres."""
    warnings.warn("Interpolating 5km geolocation to 250m resolution may result in poor quality")
    return interpolate(lon1, lat1, satz1, 5000, 250, lon1.shape[1])


def simple_modis_1km_to_250m(lon1, lat1):
    """``simple_modis_interpolator.modis_1km_to_250m`` stopped before the back-transform."""
    return _simple(lon1, lat1, 1000, 250, ecef=True)


def simple_modis_1km_to_500m(lon1, lat1):
    """``simple_modis_interpolator.modis_1km_to_500m`` stopped before the back-transform."""
    return _simple(lon1, lat1, 1000, 500, ecef=True)
