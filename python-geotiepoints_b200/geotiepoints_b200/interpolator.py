"""Generic rectilinear-grid interpolators on the GPU.

Drop-in for the modern classes of ``geotiepoints.interpolator``
(``/root/reference/geotiepoints/interpolator.py:232-378``; SURVEY.md 8f row 4, third item):
``SingleGridInterpolator`` / ``MultipleGridInterpolator`` (the reference wraps
``scipy.interpolate.RegularGridInterpolator``) and ``SingleSplineInterpolator`` / ``MultipleSplineInterpolator``
(``scipy.interpolate.RectBivariateSpline``), with the same constructor and ``interpolate`` /
``interpolate_to_shape`` arguments.  ``geointerpolator.py`` builds the lon / lat classes on them.

What is covered, per axis of the tensor product (kernels: ``csrc/gtp_grid.cu`` behind ``gtp_grid_interp_f64`` /
``gtp_geogrid_interp_f64``):

* grid classes: ``method`` "linear" (default), "nearest", "slinear", "cubic" - at construction or per call, as in
  scipy; ``bounds_error`` (default True: ``ValueError`` for fine points outside the tie-point grid),
  ``fill_value`` (default NaN; ``None`` = extrapolate).  "quintic" and "pchip" raise ``NotImplementedError``.
  "cubic" is the exact interpolating not-a-knot spline; scipy's default construction solves for it iteratively
  (``gcrotmk``, relative tolerance 1e-5), so the reference itself is only that close to it - pass
  ``solver=scipy.sparse.linalg.spsolve`` to the reference and both agree to 1e-13 (``solver`` / ``solver_args``
  are accepted and ignored here);
* spline classes: ``kx``, ``ky`` in {1, 3} (default 3), ``s = 0``; fine points outside the tie-point grid get
  the value at its edge, as FITPACK does.  Other orders, smoothing, ``bbox`` and derivative / ``grid=False``
  evaluation raise ``NotImplementedError``.

Containers: numpy arrays (host path: staged through device memory inside the C call), float64 / float32 torch
CUDA tensors (device path, results stay on the GPU).  Arithmetic is float64; results are cast to the dtype of the
tie-point values like the reference does (``.astype(self.values.dtype)``, ``:326``, ``:339``).  ``chunks=...``
builds the same dask graph as the reference (``:261-284``) and therefore needs dask.  There is no CPU fallback.
"""
import ctypes
from functools import partial

import numpy as np

from . import _capi
from ._operator import _is_torch_tensor

NEAREST, LINEAR, CUBIC = 0, 1, 3
OOB_EXTRAPOLATE, OOB_FILL, OOB_CLAMP = 0, 1, 2

_GRID_METHODS = {"nearest": NEAREST, "linear": LINEAR, "slinear": LINEAR, "cubic": CUBIC}
_ALL_GRID_METHODS = ("linear", "nearest", "slinear", "cubic", "quintic", "pchip")      # scipy's _ALL_METHODS


def _as_axis(points, dim):
    p = np.asarray(points)
    if p.ndim != 1:
        raise ValueError(f"The points in dimension {dim} must be 1-dimensional")
    p = p.astype(np.float64)
    if p.size >= 2 and np.all(p[1:] < p[:-1]):
        return p[::-1].copy(), True
    if not np.all(p[1:] > p[:-1]):
        raise ValueError(f"The points in dimension {dim} must be strictly ascending or descending")
    return np.ascontiguousarray(p), False


def _flip(values, axis):
    if _is_torch_tensor(values):
        return values.flip(axis)
    return np.flip(values, axis)


class _GridEngine:
    """Tie-point grid + one or several value arrays on it; evaluates tensor-product schemes through the C ABI."""

    def __init__(self, points, values_list):
        if len(points) != 2:
            raise ValueError(f"There are {len(points)} point arrays, but values has 2 dimensions")
        first = values_list[0]
        if len(first.shape) != 2:
            raise ValueError(f"There are 2 point arrays, but values has {len(first.shape)} dimensions")
        axes = []
        for dim, p in enumerate(points):
            axis, flipped = _as_axis(p, dim)
            if axis.size != first.shape[dim]:
                raise ValueError(f"There are {axis.size} points and {first.shape[dim]} values in dimension {dim}")
            axes.append((axis, flipped))
        self.tie = [a for a, _ in axes]
        self.on_device = _is_torch_tensor(first)
        prepared = []
        for v in values_list:
            if tuple(v.shape) != tuple(first.shape) or _is_torch_tensor(v) != self.on_device:
                raise ValueError("The tie-point arrays must have the same shape and container")
            for dim, (_, flipped) in enumerate(axes):
                if flipped:
                    v = _flip(v, dim)
            prepared.append(v)
        if self.on_device:
            import torch
            if not first.is_cuda:
                raise TypeError("torch inputs must be CUDA tensors (numpy arrays take the host path)")
            self.device = first.device
            self.values = [v.to(torch.float64).contiguous() for v in prepared]
            self.tie_dev = [torch.from_numpy(t).to(self.device) for t in self.tie]
        else:
            self.values = [np.ascontiguousarray(v, dtype=np.float64) for v in prepared]

    def check_bounds(self, fine):
        for dim, (tie, p) in enumerate(zip(self.tie, fine)):
            if p.size and not (np.all(tie[0] <= p) and np.all(p <= tie[-1])):
                raise ValueError(f"One of the requested xi is out of bounds in dimension {dim}")

    def evaluate(self, fine_points, methods, oob, fill, geo, which=None):
        """-> list of float64 arrays (geo: [lons, lats]; which = k: only array k) of shape (len(fine_y), len(fine_x))."""
        lib = _capi.lib()
        values = self.values if which is None else [self.values[which]]
        fine = [np.ascontiguousarray(np.asarray(p, dtype=np.float64).ravel()) for p in fine_points]
        for dim, m in enumerate(methods):
            if m == CUBIC and self.tie[dim].size < 4:
                raise ValueError(f"There are {self.tie[dim].size} points in dimension {dim}, but method cubic "
                                 "requires at least  4 points per dimension.")
        ny, nx, my, mx = self.tie[0].size, self.tie[1].size, fine[0].size, fine[1].size
        n = len(values)
        fill = float("nan") if fill is None else float(fill)
        if self.on_device:
            import torch
            fdev = [torch.from_numpy(f).to(self.device) for f in fine]
            stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            if geo:
                out = torch.empty((2, my, mx), dtype=torch.float64, device=self.device)
                _capi.check(lib.gtp_geogrid_interp_f64(
                    self.tie_dev[0].data_ptr(), ny, self.tie_dev[1].data_ptr(), nx, values[0].data_ptr(),
                    values[1].data_ptr(), fdev[0].data_ptr(), my, fdev[1].data_ptr(), mx, methods[0], methods[1],
                    oob, fill, out[0].data_ptr(), out[1].data_ptr(), self.device.index, stream))
                return [out[0], out[1]]
            stacked = values[0] if n == 1 else torch.stack(values)
            out = torch.empty((n, my, mx), dtype=torch.float64, device=self.device)
            _capi.check(lib.gtp_grid_interp_f64(
                self.tie_dev[0].data_ptr(), ny, self.tie_dev[1].data_ptr(), nx, stacked.data_ptr(), n,
                fdev[0].data_ptr(), my, fdev[1].data_ptr(), mx, methods[0], methods[1], oob, fill, out.data_ptr(),
                self.device.index, stream))
            return [out[k] for k in range(n)]
        if geo:
            out = np.empty((2, my, mx), np.float64)
            _capi.check(lib.gtp_geogrid_interp_host_f64(
                self.tie[0].ctypes.data, ny, self.tie[1].ctypes.data, nx, values[0].ctypes.data,
                values[1].ctypes.data, fine[0].ctypes.data, my, fine[1].ctypes.data, mx, methods[0], methods[1],
                oob, fill, out[0].ctypes.data, out[1].ctypes.data, -1))
            return [out[0], out[1]]
        stacked = values[0] if n == 1 else np.stack(values)
        out = np.empty((n, my, mx), np.float64)
        _capi.check(lib.gtp_grid_interp_host_f64(
            self.tie[0].ctypes.data, ny, self.tie[1].ctypes.data, nx, stacked.ctypes.data, n, fine[0].ctypes.data, my,
            fine[1].ctypes.data, mx, methods[0], methods[1], oob, fill, out.ctypes.data, -1))
        return [out[k] for k in range(n)]


def _cast(arr, dtype):
    """``.astype(self.values.dtype)`` of the reference for numpy and torch results."""
    if _is_torch_tensor(arr):
        import torch
        return arr.to(getattr(torch, np.dtype(dtype).name))
    return arr.astype(dtype, copy=False)


def _values_dtype(values):
    """dtype scipy keeps the values in: floating stays, everything else becomes float64."""
    if _is_torch_tensor(values):
        name = str(values.dtype).split(".")[-1]
        return np.dtype(name) if name in ("float32", "float64") else np.dtype(np.float64)
    dt = np.asarray(values).dtype
    return dt if dt in (np.float32, np.float64) else np.dtype(np.float64)


def _enumerate_chunk_slices(chunks):
    """(block position, [row slice, column slice, ...]) of every block of a normalised `chunks` tuple - the keys and
    arguments of the reference's dask graph (``interpolator.py:303-312``)."""
    starts = [np.concatenate(([0], np.cumsum(sizes))) for sizes in chunks]
    for position in np.ndindex(*(len(sizes) for sizes in chunks)):
        yield position, [slice(int(edge[k]), int(edge[k + 1])) for edge, k in zip(starts, position)]


class _GridRules:
    """RegularGridInterpolator's arguments: method, bounds_error, fill_value (+ ignored solver arguments)."""

    def __init__(self, method="linear", bounds_error=True, fill_value=np.nan, solver=None, solver_args=None):
        if method not in _ALL_GRID_METHODS:
            raise ValueError(f"Method '{method}' is not defined")
        self.method, self.bounds_error, self.fill_value = method, bounds_error, fill_value

    def resolve(self, engine, fine, method=None, nu=None):
        method = self.method if method is None else method
        if method not in _ALL_GRID_METHODS:
            raise ValueError(f"Method '{method}' is not defined")
        if method not in _GRID_METHODS:
            raise NotImplementedError(f"method '{method}' is not available on the GPU (linear, nearest, slinear, cubic are)")
        if nu is not None:
            raise NotImplementedError("derivatives (nu) are not available on the GPU")
        if self.bounds_error:
            engine.check_bounds(fine)
        oob = OOB_EXTRAPOLATE if (self.bounds_error or self.fill_value is None) else OOB_FILL
        return (_GRID_METHODS[method],) * 2, oob, self.fill_value


class _SplineRules:
    """RectBivariateSpline's arguments: kx (rows), ky (columns), s, bbox."""

    def __init__(self, bbox=None, kx=3, ky=3, s=0, maxit=20):
        if bbox is not None and any(b is not None for b in bbox):
            raise NotImplementedError("bbox is not available on the GPU")
        if s != 0:
            raise NotImplementedError("smoothing splines (s != 0) are not available on the GPU")
        for k in (kx, ky):
            if k not in (1, 3):
                raise NotImplementedError("spline orders 1 and 3 are available on the GPU")
        self.methods = (LINEAR if kx == 1 else CUBIC, LINEAR if ky == 1 else CUBIC)

    def resolve(self, engine, fine, dx=0, dy=0, grid=True):
        if dx or dy or not grid:
            raise NotImplementedError("derivatives and grid=False are not available on the GPU")
        for name, p in zip("xy", fine):
            if p.size > 1 and not np.all(p[1:] > p[:-1]):
                raise ValueError(f"{name} must be strictly increasing when `grid` is True")     # RectBivariateSpline's
        return self.methods, OOB_CLAMP, None


class AbstractSingleInterpolator:
    """An interpolator for a single 2d data array (``interpolator.py:232-300``)."""

    _geo = False

    def __init__(self, points, values, rules, engine=None, index=0):
        self.points = points
        self.values = values
        self._rules = rules
        self._dtype = _values_dtype(values)
        self._engine = engine if engine is not None else _GridEngine(points, [values])       # shared by a Multiple...
        self._index = index

    def interpolate(self, fine_points, chunks=None, **interpolator_call_kwargs):
        """Interpolate the value points to the *fine_points* grid (``:245-259``)."""
        if chunks is not None:
            return self.interpolate_dask(fine_points, chunks=chunks, **interpolator_call_kwargs)
        return self.interpolate_numpy(fine_points, **interpolator_call_kwargs)

    def interpolate_dask(self, fine_points, chunks, **interpolator_call_kwargs):
        """Interpolate (lazily) to a dask array (``:261-284``)."""
        from dask.base import tokenize
        import dask.array as da
        from dask.array.core import normalize_chunks
        v_fine_points, h_fine_points = fine_points
        shape = len(v_fine_points), len(h_fine_points)
        chunks = normalize_chunks(chunks, shape, dtype=self._dtype)
        token = tokenize(chunks, self.points, self.values, fine_points, interpolator_call_kwargs)
        name = 'interpolate-' + token
        interpolate_slices = partial(self.interpolate_slices, **interpolator_call_kwargs)
        dskx = {(name, ) + position: (interpolate_slices, slices)
                for position, slices in _enumerate_chunk_slices(chunks)}
        return da.Array(dskx, name, shape=list(shape), chunks=chunks, dtype=self._dtype)

    def interpolate_numpy(self, fine_points, **interpolator_call_kwargs):
        """Interpolate to an array of the tie points' container (numpy, or a torch CUDA tensor)."""
        fine = [np.asarray(p, dtype=np.float64).ravel() for p in fine_points]
        methods, oob, fill = self._rules.resolve(self._engine, fine, **interpolator_call_kwargs)
        out = self._engine.evaluate(fine, methods, oob, fill, geo=False, which=self._index)
        return _cast(out[0], self._dtype)

    def interpolate_slices(self, fine_points, **interpolator_call_kwargs):
        """Interpolate using slices: *fine_points* are a tuple of slices for the y and x dimensions (``:290-300``)."""
        slice_y, slice_x = fine_points
        points_y = np.arange(slice_y.start, slice_y.stop)
        points_x = np.arange(slice_x.start, slice_x.stop)
        return self.interpolate_numpy((points_y, points_x), **interpolator_call_kwargs)


class SingleGridInterpolator(AbstractSingleInterpolator):
    """A regular grid interpolator for a single 2d data array (``:315-325``)."""

    def __init__(self, points, values, _engine=None, _index=0, **interpolator_init_kwargs):
        super().__init__(points, values, _GridRules(**interpolator_init_kwargs), _engine, _index)


class SingleSplineInterpolator(AbstractSingleInterpolator):
    """A spline interpolator for a single 2d data array (``:328-339``)."""

    def __init__(self, points, values, _engine=None, _index=0, **interpolator_init_kwargs):
        super().__init__(points, values, _SplineRules(**interpolator_init_kwargs), _engine, _index)


class AbstractMultipleInterpolator:
    """Interpolator that works on multiple arrays (``:342-361``): one launch for all of them."""

    _single = None

    def __init__(self, tie_points, *data, **interpolator_init_kwargs):
        self._engine = _GridEngine(tie_points, list(data)) if data else None        # one upload for all arrays
        self.interpolators = [self._single(tie_points, values, _engine=self._engine, _index=k, **interpolator_init_kwargs)
                              for k, values in enumerate(data)]
        self._rules = self.interpolators[0]._rules if data else None

    def interpolate(self, fine_points, chunks=None, **kwargs):
        """Interpolate the data; keyword arguments as for the single interpolators' ``interpolate``."""
        if chunks is not None:
            return (interpolator.interpolate(fine_points, chunks=chunks, **kwargs) for interpolator in self.interpolators)
        fine = [np.asarray(p, dtype=np.float64).ravel() for p in fine_points]
        methods, oob, fill = self._rules.resolve(self._engine, fine, **kwargs)
        out = self._engine.evaluate(fine, methods, oob, fill, geo=False)
        return (_cast(o, single._dtype) for o, single in zip(out, self.interpolators))

    def interpolate_to_shape(self, shape, **interpolator_call_kwargs):
        """Interpolate to a given *shape* (``:358-361``)."""
        fine_points = [np.arange(size) for size in shape]
        return self.interpolate(fine_points, **interpolator_call_kwargs)


class MultipleGridInterpolator(AbstractMultipleInterpolator):
    """Grid interpolator that works on multiple data arrays (``:364-369``)."""

    _single = SingleGridInterpolator


class MultipleSplineInterpolator(AbstractMultipleInterpolator):
    """Spline interpolator that works on multiple data arrays (``:372-377``)."""

    _single = SingleSplineInterpolator
