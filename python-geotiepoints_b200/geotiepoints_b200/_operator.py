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


def empty_outputs(n_out, out_shape, dtype):
    """``n_out`` numpy outputs as the planes of ONE page-locked ``(n_out, rows, cols)`` block: the tuple the
    reference's operators return, and - for the dask adapter, which needs them stacked
    (``_modis_utils.pyx:191-198``) - a single array again without a copy (:func:`stack`)."""
    block = _capi.pinned_pool.empty((n_out,) + tuple(out_shape), dtype)
    return tuple(block[k] for k in range(n_out))


def empty_outputs_torch(n_out, out_shape, dtype, device):
    import torch
    block = torch.empty((n_out,) + tuple(out_shape), dtype=dtype, device=device)
    return tuple(block[k] for k in range(n_out))


def stack(outputs):
    """``(n, rows, cols)`` array of an operator's outputs.  Zero copy when they are the adjacent planes of one
    block, which is what :func:`run` returns; any other arrays are concatenated (numpy) / stacked (torch)."""
    first = outputs[0]
    n = len(outputs)
    if _is_torch_tensor(first):
        import torch
        same = all(o.is_contiguous() and o.shape == first.shape and o.dtype == first.dtype
                   and o.untyped_storage().data_ptr() == first.untyped_storage().data_ptr()
                   and o.data_ptr() == first.data_ptr() + k * first.numel() * first.element_size()
                   for k, o in enumerate(outputs))
        if same:
            return torch.as_strided(first, (n,) + tuple(first.shape), (first.numel(),) + tuple(first.stride()))
        return torch.stack(list(outputs))
    outputs = [np.asarray(o) for o in outputs]
    first = outputs[0]
    same = all(o.flags.c_contiguous and o.shape == first.shape and o.dtype == first.dtype
               and o.base is not None and o.base is first.base
               and o.ctypes.data == first.ctypes.data + k * first.nbytes for k, o in enumerate(outputs))
    if same:
        return np.lib.stride_tricks.as_strided(first, shape=(n,) + first.shape, strides=(first.nbytes,) + first.strides)
    return np.concatenate([o[np.newaxis] for o in outputs], axis=0)


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
    out = empty_outputs(3 if ecef else 2, out_shape, arrays[0].dtype)
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
    out = empty_outputs_torch(3 if ecef else 2, out_shape, first.dtype, first.device)
    stream = torch.cuda.current_stream(first.device).cuda_stream
    fn = getattr(lib, f"gtp_{kind}_interp_{'xyz_' if ecef else ''}{suffix}")
    rc = fn(*[a.data_ptr() for a in arrays], n_scans, *scalar_args,
            *[o.data_ptr() for o in out], first.device.index, ctypes.c_void_p(stream))
    _capi.check(rc)
    return tuple(out)
