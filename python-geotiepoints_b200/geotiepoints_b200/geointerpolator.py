"""Geographical (lon / lat) grid interpolation on the GPU.

Drop-in for ``geotiepoints.geointerpolator.GeoGridInterpolator`` and ``GeoSplineInterpolator``
(``/root/reference/geotiepoints/geointerpolator.py:76-113``; SURVEY.md 8f row 4, third item): the grid / spline
interpolators of ``interpolator.py`` run on the Cartesian x / y / z of the tie points (``lonlat2xyz``, ``:52-60``,
sphere of 6370997 m) and the result is turned back into lon / lat (``xyz2lonlat``, ``:63-75``).  Here that is one
fused call (``gtp_geogrid_interp_f64``, ``csrc/gtp_grid.cu``): conversion of the tie points, the tensor-product
scheme and an anchor-relative back-transform; the Cartesian planes are never written out.  Constructor and
``interpolate`` / ``interpolate_to_shape`` arguments are those of the reference (see ``interpolator.py`` for which
scipy options are covered).  The legacy ``GeoInterpolator`` class is served through the package-level
``modis5kmto1km`` / ``modis1kmto500m`` / ``modis1kmto250m`` (``legacy.py``) only.

Tie points: ``(lons, lats)`` arrays (numpy, or torch CUDA tensors), or one pyresample-like definition object
(``.lons`` / ``.lats``).  Results are cast to the dtype the reference's ``lonlat2xyz`` gives (float32 stays
float32, everything else is float64); the arithmetic is float64 throughout, so for float32 tie points the
reference's float32 rounding of the Cartesian tie points and of its back-transform is not reproduced (results
agree to a float32 ulp).  There is no CPU fallback.
"""
import numpy as np

from . import interpolator as _itp
from ._operator import _is_torch_tensor

EARTH_RADIUS = 6370997.0


def _lonlat_arrays(data):
    """to_xyz (``geointerpolator.py:95-108``) without the conversion: the kernel does it."""
    if len(data) == 1:
        swath = data[0]
        if not (hasattr(swath, "lons") and hasattr(swath, "lats")):
            raise TypeError("a geometry definition must provide .lons and .lats")
        lons, lats = swath.lons, swath.lats
        lons, lats = getattr(lons, "values", lons), getattr(lats, "values", lats)
    elif len(data) == 2:
        lons, lats = data
    else:
        raise ValueError("Either pass lon/lats or a pyresample definition.")
    if not _is_torch_tensor(lons):
        lons, lats = np.asarray(lons), np.asarray(lats)
    return lons, lats


def _work_with_lonlats(single, name):
    """Adapt the interpolator classes to geographical coordinates (``geointerpolator.py:76-92``)."""

    class GeoKlass:

        def __init__(self, tie_points, *data, **interpolator_init_kwargs):
            """Set up the interpolator."""
            lons, lats = _lonlat_arrays(data)
            self._engine = _itp._GridEngine(tie_points, [lons, lats])
            self._probe = single(tie_points, lons, _engine=self._engine, **interpolator_init_kwargs)
            self._rules = self._probe._rules
            self._dtype = _itp._values_dtype(lons)
            self.tie_points = tie_points

        def interpolate(self, fine_points, chunks=None, **interpolator_call_kwargs):
            """Interpolate to *fine_points*: (lons, lats)."""
            if chunks is not None:
                return self._interpolate_dask(fine_points, chunks, **interpolator_call_kwargs)
            return self._interpolate_now(fine_points, **interpolator_call_kwargs)

        def interpolate_to_shape(self, shape, **interpolator_call_kwargs):
            """Interpolate to a given *shape*."""
            return self.interpolate([np.arange(size) for size in shape], **interpolator_call_kwargs)

        def _interpolate_now(self, fine_points, **interpolator_call_kwargs):
            fine = [np.asarray(p, dtype=np.float64).ravel() for p in fine_points]
            methods, oob, fill = self._rules.resolve(self._engine, fine, **interpolator_call_kwargs)
            lons, lats = self._engine.evaluate(fine, methods, oob, fill, geo=True)
            return _itp._cast(lons, self._dtype), _itp._cast(lats, self._dtype)

        def _interpolate_slices(self, slices, **interpolator_call_kwargs):
            _, slice_y, slice_x = slices
            fine = np.arange(slice_y.start, slice_y.stop), np.arange(slice_x.start, slice_x.stop)
            lons, lats = self._interpolate_now(fine, **interpolator_call_kwargs)
            if _is_torch_tensor(lons):
                import torch
                return torch.stack((lons, lats))
            return np.stack((lons, lats))

        def _interpolate_dask(self, fine_points, chunks, **interpolator_call_kwargs):
            """The reference's graph (``interpolator.py:261-284``) with lon and lat of a chunk from one launch."""
            from functools import partial
            from dask.base import tokenize
            import dask.array as da
            from dask.array.core import normalize_chunks
            v_fine_points, h_fine_points = fine_points
            shape = len(v_fine_points), len(h_fine_points)
            chunks = normalize_chunks(chunks, shape, dtype=self._dtype)
            token = tokenize(chunks, self.tie_points, fine_points, interpolator_call_kwargs, id(self))
            name = 'geo-interpolate-' + token
            func = partial(self._interpolate_slices, **interpolator_call_kwargs)
            chunks3 = ((2,),) + tuple(chunks)
            dskx = {(name,) + position: (func, slices) for position, slices in _itp._enumerate_chunk_slices(chunks3)}
            both = da.Array(dskx, name, shape=[2] + list(shape), chunks=chunks3, dtype=self._dtype)
            return both[0], both[1]

    GeoKlass.__name__ = GeoKlass.__qualname__ = name
    return GeoKlass


GeoGridInterpolator = _work_with_lonlats(_itp.SingleGridInterpolator, "GeoGridInterpolator")
GeoSplineInterpolator = _work_with_lonlats(_itp.SingleSplineInterpolator, "GeoSplineInterpolator")
