"""Common host plumbing of the two operators: validation, buffers, C-ABI dispatch."""
import ctypes

import numpy as np

from . import _capi


def _is_torch_tensor(obj):
    return type(obj).__module__.startswith("torch") and hasattr(obj, "data_ptr")


def _check_same_dtype(arrays):
    """The errors Cython's typed-buffer arguments raise."""
    first = arrays[0].dtype
    suffix = _capi.suffix(np.dtype(str(first).replace("torch.", "")))
    for other in arrays[1:]:
        if other.dtype != first:
            want = "float" if suffix == "f32" else "double"
            raise ValueError(f"Buffer dtype mismatch, expected '{want}' but got a different dtype "
                             f"({other.dtype}); all inputs must share one dtype")
    return suffix


def _check_2d_same_shape(arrays):
    for a in arrays:
        if a.ndim != 2:
            raise ValueError("Buffer has wrong number of dimensions (expected 2, got %d)" % a.ndim)
    shape = tuple(arrays[0].shape)
    for a in arrays[1:]:
        if tuple(a.shape) != shape:
            raise ValueError(f"input arrays must have the same shape, got {shape} and {tuple(a.shape)}")
    return shape


def run(kind, arrays, n_scans, out_shape, scalar_args, ecef=False):
    """Run one operator.

    ``kind`` is ``"cviirs"`` or ``"simple"``; ``arrays`` are the 2 or 3 inputs (all numpy
    or all torch CUDA tensors); ``scalar_args`` are the integer arguments between
    ``n_scans`` and the output pointers in the C signature.
    Returns ``(lon, lat)`` - with ``ecef`` the Cartesian ``(x, y, z)`` the reference back-transforms
    (``gtp_*_interp_xyz_*``) - of the same container type and dtype as the inputs.
    """
    lib = _capi.lib()
    if all(_is_torch_tensor(a) for a in arrays):
        return _run_torch(lib, kind, arrays, n_scans, out_shape, scalar_args, ecef)
    arrays = [np.ascontiguousarray(a) for a in arrays]
    suffix = _check_same_dtype(arrays)
    out = [_capi.pinned_pool.empty(out_shape, arrays[0].dtype) for _ in range(3 if ecef else 2)]
    fn = getattr(lib, f"gtp_{kind}_interp_{'xyz_' if ecef else ''}host_{suffix}")
    ptrs = [a.ctypes.data for a in arrays]
    rc = fn(*ptrs, n_scans, *scalar_args, *[o.ctypes.data for o in out], -1)
    _capi.check(rc)
    return tuple(out)


def _run_torch(lib, kind, arrays, n_scans, out_shape, scalar_args, ecef=False):
    import torch
    first = arrays[0]
    if not first.is_cuda:
        raise TypeError("torch inputs must be CUDA tensors (numpy arrays take the host path)")
    arrays = [a.contiguous() for a in arrays]
    suffix = _check_same_dtype(arrays)
    for a in arrays[1:]:
        if a.device != first.device:
            raise ValueError("all input tensors must live on the same device")
    out = [torch.empty(out_shape, dtype=first.dtype, device=first.device) for _ in range(3 if ecef else 2)]
    stream = torch.cuda.current_stream(first.device).cuda_stream
    fn = getattr(lib, f"gtp_{kind}_interp_{'xyz_' if ecef else ''}{suffix}")
    rc = fn(*[a.data_ptr() for a in arrays], n_scans, *scalar_args,
            *[o.data_ptr() for o in out], first.device.index, ctypes.c_void_p(stream))
    _capi.check(rc)
    return tuple(out)
