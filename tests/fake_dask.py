"""A tiny stand-in for ``dask.array`` (dask is not installed in the image).

It models exactly what ``scanline_mapblocks`` touches: ``chunks`` / ``shape`` / ``ndim`` /
``dtype``, ``rechunk``, ``map_blocks`` with ``new_axis=[0]`` and leading-axis indexing, and
a ``compute`` that counts how often it was triggered - the role of the reference's
``CustomScheduler`` (geotiepoints/tests/utils.py:4-28).
"""
import numpy as np

COMPUTES = {"n": 0}


def _normalize(chunks, shape):
    out = []
    for c, n in zip(chunks, shape):
        if isinstance(c, (tuple, list)):
            out.append(tuple(c))
        elif c == -1:
            out.append((n,))
        else:
            full, rest = divmod(n, c)
            out.append((c,) * full + ((rest,) if rest else ()))
    return tuple(out)


class Array:
    def __init__(self, thunk, shape, chunks, dtype):
        self._thunk, self.shape, self.dtype = thunk, tuple(shape), np.dtype(dtype)
        self.chunks = _normalize(chunks, shape)
        self.ndim = len(self.shape)

    def rechunk(self, chunks):
        return Array(self._thunk, self.shape, chunks, self.dtype)

    def compute(self):
        COMPUTES["n"] += 1
        return self._thunk(self)

    def __getitem__(self, idx):
        assert isinstance(idx, (int, np.integer))
        parent = self
        return Array(lambda _self: parent._thunk(parent)[idx], self.shape[1:], self.chunks[1:], self.dtype)


def _is_torch(obj):
    return type(obj).__module__.startswith("torch")


def from_array(arr, chunks):
    """numpy arrays, or torch tensors (blocks then stay torch tensors: the GPU-resident block path)."""
    if not _is_torch(arr):
        arr = np.asarray(arr)
    dtype = np.dtype(str(arr.dtype).replace("torch.", ""))
    return Array(lambda _self: arr, tuple(arr.shape), (chunks,) * arr.ndim if isinstance(chunks, int) else chunks, dtype)


BLOCK_SHAPES = []


def map_blocks(func, *arrays, new_axis=None, chunks=None, dtype=None, meta=None, **kwargs):
    assert new_axis == [0]
    first = arrays[0]
    chunks = tuple((c,) if isinstance(c, int) else tuple(c) for c in chunks)
    out_shape = tuple(sum(c) for c in chunks)

    def thunk(_self):
        data = [a._thunk(a) for a in arrays]
        row_chunks = first.chunks[0]
        assert all(a.chunks == first.chunks for a in arrays), "map_blocks needs aligned chunks"
        assert len(first.chunks[1]) == 1, "blocks must span the full width"
        pieces, r0 = [], 0
        for rc in row_chunks:
            block = [d[r0:r0 + rc] for d in data]
            BLOCK_SHAPES.append(block[0].shape)
            pieces.append(func(*block, **kwargs))
            r0 += rc
        if _is_torch(pieces[0]):
            import torch
            return torch.cat(pieces, dim=1)
        return np.concatenate(pieces, axis=1)

    return Array(thunk, out_shape, chunks, dtype)


def compute(*arrays):
    return tuple(a.compute() for a in arrays)
