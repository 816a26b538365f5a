"""Scan-aligned array-container adapter (host side of the drop-in boundary).

Re-hosts, in plain Python, the behaviour of the decorator that the reference keeps
inside a Cython module (``/root/reference/geotiepoints/_modis_utils.pyx:80-198``):
numpy inputs go straight to the operator; dask arrays (bare or inside xarray
``DataArray``s) are rechunked to whole scans x full width and mapped block by block,
lazily, with the results re-wrapped in ``DataArray``s when the inputs were.  Error
messages and chunk arithmetic are the reference's.  torch CUDA tensors are also passed
straight through (zero-copy device path).
"""
import functools

import numpy as np

try:
    import dask.array as da
except ImportError:  # dask is optional, exactly as in the reference (:11-15)
    da = None

try:
    import xarray as xr
except ImportError:  # (:17-20)
    xr = None

_ROWS_PER_SCAN = {5000: 2, 1000: 10, 500: 20, 250: 40}
_FINE_PIXELS_PER_1KM = {250: 4, 500: 2, 1000: 1}


def rows_per_scan_for_resolution(res):
    """Rows of one MODIS scan at ``res`` metres (``_modis_utils.pyx:80-86``)."""
    return _ROWS_PER_SCAN[res]


def _is_array(obj):
    return hasattr(obj, "ndim")


def _is_lazy(obj):
    return hasattr(obj, "compute")


def scanline_mapblocks(func=None, *, n_outputs=2):
    """Decorator: call ``func`` per whole-scan block for dask inputs (``:89-129``).

    ``n_outputs``: how many same-shaped arrays ``func`` returns (2 = lon / lat as in the reference; the ECEF
    operators return 3).  A keyword argument ``out_dtype`` of ``func``, when given, is the dtype of the
    results (``mod03``); otherwise they have the dtype of the first input, as in the reference (``:150``).
    """
    if func is None:
        return functools.partial(scanline_mapblocks, n_outputs=n_outputs)

    @functools.wraps(func)
    def adapter(*args, coarse_resolution=None, fine_resolution=None, **kwargs):
        if coarse_resolution is None or fine_resolution is None:
            raise ValueError("'coarse_resolution' and 'fine_resolution' are required keyword arguments.")
        template = next(a for a in args if _is_array(a))
        if template.ndim != 2:
            raise ValueError("Expected 2D input arrays.")
        if not _is_lazy(template):
            return func(*args, coarse_resolution=coarse_resolution, fine_resolution=fine_resolution, **kwargs)

        blocks = [a.data if hasattr(a, "dims") else a for a in args]
        blocks = _align_chunks_to_scans(blocks, rows_per_scan_for_resolution(coarse_resolution))
        lazy = _map_scan_blocks(func, n_outputs, coarse_resolution, fine_resolution, *blocks, **kwargs)
        if hasattr(template, "dims"):
            return [xr.DataArray(res, dims=template.dims) for res in lazy if isinstance(res, da.Array)]
        return lazy

    return adapter


def _align_chunks_to_scans(arrays, rows_per_scan):
    """Rechunk so that every block is whole scans x the full width (``:165-188``)."""
    template = next(a for a in arrays if _is_array(a))
    row_chunks, col_chunks = template.chunks
    n_rows, n_cols = template.shape
    if n_rows % rows_per_scan != 0:
        raise ValueError("Input longitude/latitude data does not consist of "
                         "whole scans (10 rows per scan).")
    chunked = [a.chunks for a in arrays if hasattr(a, "chunks")]
    rows_ok = all(c % rows_per_scan == 0 for c in row_chunks)
    # the reference's column test, kept verbatim in meaning (:173): a single column
    # chunk that is *not* the full width counts as "good"; anything else rechunks.
    cols_ok = len(col_chunks) == 1 and col_chunks[0] != n_cols
    same = all(chunked[0] == other for other in chunked[1:])
    if rows_ok and cols_ok and same:
        return arrays
    rows = (row_chunks[0] // rows_per_scan) * rows_per_scan
    return [a.rechunk((rows, -1)) if hasattr(a, "chunks") else a for a in arrays]


def _map_scan_blocks(func, n_outputs, coarse_resolution, fine_resolution, *arrays, **kwargs):
    """``da.map_blocks`` with a leading (lon, lat) axis of length 2 (``:136-152``).

    The operator writes its outputs into the planes of one page-locked (numpy blocks) or device (torch
    blocks) ``(n, rows, cols)`` buffer, so the stacking the reference does with ``np.concatenate``
    (``:191-198``) is a view here, and blocks that live on the GPU stay there."""
    from ._operator import stack
    template = next(a for a in arrays if _is_array(a))
    factor = coarse_resolution // fine_resolution
    out_rows = tuple(c * factor for c in template.chunks[0])
    out_cols = (1354 * _FINE_PIXELS_PER_1KM[fine_resolution],)
    out_dtype = np.dtype(kwargs.get("out_dtype") or template.dtype)

    @functools.wraps(func)
    def stacked(*block_args, **block_kwargs):
        return stack(func(*block_args, **block_kwargs))

    res = da.map_blocks(stacked, *arrays,
                        coarse_resolution=coarse_resolution, fine_resolution=fine_resolution, **kwargs,
                        new_axis=[0], chunks=(n_outputs, out_rows, out_cols), dtype=out_dtype,
                        meta=np.empty((2, 2, 2), dtype=out_dtype))
    return tuple(res[k] for k in range(res.shape[0]))
