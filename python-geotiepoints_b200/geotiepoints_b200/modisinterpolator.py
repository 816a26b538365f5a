"""Satellite-zenith driven ("CVIIRS") interpolation of MODIS geolocation - public entry points.

Drop-in for ``geotiepoints.modisinterpolator`` (``/root/reference/geotiepoints/modisinterpolator.py:21-62``):
the same five functions ``modis_<coarse>_to_<fine>(lon1, lat1, satz1)`` returning ``(lons, lats)``,
with the same "poor quality" warning for 5 km -> 500 m / 250 m (``:45-47``, ``:56-58``).  The five
differ only in a row of the table below, so they are generated from it; each one hands its arrays to
the operator ``_modis_interpolator.interpolate``, which launches the CUDA kernels.
"""
import warnings

from ._modis_interpolator import interpolate

#: name -> (coarse resolution, fine resolution) in metres
_CONFIGURATIONS = {
    "modis_1km_to_250m": (1000, 250),
    "modis_1km_to_500m": (1000, 500),
    "modis_5km_to_1km": (5000, 1000),
    "modis_5km_to_500m": (5000, 500),
    "modis_5km_to_250m": (5000, 250),
}


def _label(metres):
    return f"{metres // 1000}km" if metres >= 1000 else f"{metres}m"


def _entry_point(name, coarse, fine):
    poor_quality = coarse == 5000 and fine < 1000

    def entry(lon1, lat1, satz1):
        extra = {}
        if coarse == 5000:
            # 5 km granules come 270 or 271 tie points wide; the operator needs to know which
            extra["coarse_scan_width"] = lon1.shape[1]
        if poor_quality:
            warnings.warn(f"Interpolating 5km geolocation to {_label(fine)} resolution may result in poor quality")
        return interpolate(lon1, lat1, satz1, coarse_resolution=coarse, fine_resolution=fine, **extra)

    entry.__name__ = entry.__qualname__ = name
    entry.__doc__ = f"Interpolate MODIS geolocation from {_label(coarse)} to {_label(fine)} resolution."
    return entry


for _name, (_coarse, _fine) in _CONFIGURATIONS.items():
    globals()[_name] = _entry_point(_name, _coarse, _fine)
del _name, _coarse, _fine

__all__ = sorted(_CONFIGURATIONS)
