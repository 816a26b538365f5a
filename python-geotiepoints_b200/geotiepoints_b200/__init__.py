"""geotiepoints_b200 - B200-native MODIS geolocation tie-point interpolation.

Drop-in for the hot path of pytroll/python-geotiepoints (``geotiepoints.modisinterpolator``
and ``geotiepoints.simple_modis_interpolator``): same function names, arguments, return
values, warnings and errors; the per-scan numeric kernels run as hand-written sm_100a
CUDA behind the C ABI declared in ``include/gtp_b200.h``.  There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import ecef, geointerpolator, interpolator, legacy, mod03, modisinterpolator, simple_modis_interpolator, viiinterpolator  # noqa: F401
from .legacy import get_scene_splits, modis1kmto250m, modis1kmto500m, modis5kmto1km  # noqa: F401  (geotiepoints/__init__.py)
from ._capi import trim  # noqa: F401
