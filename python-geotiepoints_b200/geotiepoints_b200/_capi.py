"""ctypes binding of ``libgtp_b200.so`` (C ABI: ``include/gtp_b200.h``).

The library is built in-tree by ``__graft_entry__.build()`` /
``python -m geotiepoints_b200.build``.  Loading fails loudly when it is missing and every
compute call fails loudly when there is no CUDA device: there is no CPU fallback.
"""
import ctypes
import os
import threading
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
#: the in-tree build.  GTP_B200_LIB points at an A/B build of the same ABI (tools/build_variant.sh); using it
#: is never silent: a warning at load time, and bench.py prints the path and sha256 of what it timed.
DEFAULT_LIB_PATH = os.path.join(_HERE, "libgtp_b200.so")
LIB_PATH = os.environ.get("GTP_B200_LIB") or DEFAULT_LIB_PATH

GTP_E_WIDTH_5KM = 2
GTP_E_NO_DEVICE = 5

_lib = None
_lock = threading.Lock()

_c_f32p = ctypes.POINTER(ctypes.c_float)
_c_f64p = ctypes.POINTER(ctypes.c_double)
_i64, _int, _vp = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p

#: every symbol include/gtp_b200.h declares -> (restype, argtypes)
SIGNATURES = {
    "gtp_version": (ctypes.c_char_p, []),
    "gtp_last_error": (ctypes.c_char_p, []),
    "gtp_device_count": (_int, []),
    "gtp_launch_count": (ctypes.c_uint64, []),
    "gtp_cviirs_geometry": (_int, [_int, _int, _int] + [ctypes.POINTER(_int)] * 4),
    "gtp_host_alloc": (_int, [ctypes.POINTER(_vp), ctypes.c_uint64]),
    "gtp_host_free": (_int, [_vp]),
    "gtp_trim": (_int, []),
    "gtp_set_d2h_relay": (_int, [_int]),
    "gtp_get_d2h_relay": (_int, []),
}
for _suf in ("f32", "f64"):
    SIGNATURES[f"gtp_cviirs_interp_{_suf}"] = (_int, [_vp, _vp, _vp, _i64, _int, _int, _int, _vp, _vp, _int, _vp])
    SIGNATURES[f"gtp_simple_interp_{_suf}"] = (_int, [_vp, _vp, _i64, _i64, _int, _vp, _vp, _int, _vp])
    SIGNATURES[f"gtp_cviirs_interp_host_{_suf}"] = (_int, [_vp, _vp, _vp, _i64, _int, _int, _int, _vp, _vp, _int])
    SIGNATURES[f"gtp_simple_interp_host_{_suf}"] = (_int, [_vp, _vp, _i64, _i64, _int, _vp, _vp, _int])
    # ECEF output: three output pointers instead of two
    SIGNATURES[f"gtp_cviirs_interp_xyz_{_suf}"] = (_int, [_vp, _vp, _vp, _i64, _int, _int, _int, _vp, _vp, _vp, _int, _vp])
    SIGNATURES[f"gtp_simple_interp_xyz_{_suf}"] = (_int, [_vp, _vp, _i64, _i64, _int, _vp, _vp, _vp, _int, _vp])
    SIGNATURES[f"gtp_cviirs_interp_xyz_host_{_suf}"] = (_int, [_vp, _vp, _vp, _i64, _int, _int, _int, _vp, _vp, _vp, _int])
    SIGNATURES[f"gtp_simple_interp_xyz_host_{_suf}"] = (_int, [_vp, _vp, _i64, _i64, _int, _vp, _vp, _vp, _int])
_dbl = ctypes.c_double
SIGNATURES["gtp_cviirs_interp_xyz_generic_f64"] = SIGNATURES["gtp_cviirs_interp_xyz_f64"]
SIGNATURES["gtp_simple_interp_xyz_generic_f64"] = SIGNATURES["gtp_simple_interp_xyz_f64"]
SIGNATURES["gtp_cviirs_interp_xyz_mixed"] = (_int, [_vp, _vp, _vp, _int, _dbl, _dbl, _i64, _int, _int, _vp, _vp, _vp, _int, _int, _vp])
SIGNATURES["gtp_cviirs_interp_xyz_mixed_host"] = (_int, [_vp, _vp, _vp, _int, _dbl, _dbl, _i64, _int, _int, _vp, _vp, _vp, _int, _int])
SIGNATURES["gtp_simple_interp_xyz_mixed"] = (_int, [_vp, _vp, _i64, _int, _vp, _vp, _vp, _int, _int, _vp])
SIGNATURES["gtp_simple_interp_xyz_mixed_host"] = (_int, [_vp, _vp, _i64, _int, _vp, _vp, _vp, _int, _int])
SIGNATURES["gtp_cviirs_interp_mixed"] = (_int, [_vp, _vp, _vp, _int, _dbl, _dbl, _i64, _int, _int, _int, _vp, _vp, _int, _int, _vp])
SIGNATURES["gtp_cviirs_interp_mixed_host"] = (_int, [_vp, _vp, _vp, _int, _dbl, _dbl, _i64, _int, _int, _int, _vp, _vp, _int, _int])
SIGNATURES["gtp_simple_interp_mixed"] = (_int, [_vp, _vp, _i64, _int, _vp, _vp, _int, _int, _vp])
SIGNATURES["gtp_simple_interp_mixed_host"] = (_int, [_vp, _vp, _i64, _int, _vp, _vp, _int, _int])
SIGNATURES["gtp_cviirs_interp_generic_f64"] = SIGNATURES["gtp_cviirs_interp_f64"]
SIGNATURES["gtp_simple_interp_generic_f64"] = SIGNATURES["gtp_simple_interp_f64"]
SIGNATURES["gtp_cviirs_interp_generic_f32"] = SIGNATURES["gtp_cviirs_interp_f32"]
SIGNATURES["gtp_legacy_interp_f64"] = (_int, [_vp, _vp, _i64, _int, _vp, _vp, _int, _vp])
SIGNATURES["gtp_legacy_interp_host_f64"] = (_int, [_vp, _vp, _i64, _int, _vp, _vp, _int])
SIGNATURES["gtp_vii_interp_f64"] = (_int, [_vp, _i64, _i64, _int, _int, _vp, _int, _vp])
SIGNATURES["gtp_vii_interp_host_f64"] = (_int, [_vp, _i64, _i64, _int, _int, _vp, _int])
SIGNATURES["gtp_vii_geo_interp_f64"] = (_int, [_vp, _vp, _i64, _i64, _int, _int, _int, _dbl, _vp, _vp, _int, _vp])
SIGNATURES["gtp_grid_interp_f64"] = (_int, [_vp, _i64, _vp, _i64, _vp, _int, _vp, _i64, _vp, _i64, _int, _int, _int, _dbl, _vp, _int, _vp])
SIGNATURES["gtp_grid_interp_host_f64"] = (_int, [_vp, _i64, _vp, _i64, _vp, _int, _vp, _i64, _vp, _i64, _int, _int, _int, _dbl, _vp, _int])
SIGNATURES["gtp_geogrid_interp_f64"] = (_int, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _int, _int, _int, _dbl, _vp, _vp, _int, _vp])
SIGNATURES["gtp_geogrid_interp_host_f64"] = (_int, [_vp, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _int, _int, _int, _dbl, _vp, _vp, _int])
SIGNATURES["gtp_vii_geo_interp_host_f64"] = (_int, [_vp, _vp, _i64, _i64, _int, _int, _int, _dbl, _vp, _vp, _int])


class GtpError(RuntimeError):
    """A failure reported by libgtp_b200 (CUDA error or invalid argument)."""


def lib():
    """Load (once) and return the ctypes handle; raise if the CUDA library is not built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(
                        f"{LIB_PATH} is missing: build the sm_100a CUDA library first "
                        "(python -m geotiepoints_b200.build).  geotiepoints_b200 has no CPU fallback.")
                if os.path.realpath(LIB_PATH) != os.path.realpath(DEFAULT_LIB_PATH):
                    import warnings
                    warnings.warn(f"geotiepoints_b200: GTP_B200_LIB is set, loading {LIB_PATH} instead of the in-tree "
                                  f"library", RuntimeWarning, stacklevel=2)
                handle = ctypes.CDLL(LIB_PATH)
                for name, (restype, argtypes) in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype, fn.argtypes = restype, argtypes
                _lib = handle
    return _lib


def library_info():
    """Path, sha256 and version string of the loaded library (bench.py prints them with every line)."""
    import hashlib
    h = hashlib.sha256()
    with open(LIB_PATH, "rb") as fh:
        for block in iter(lambda: fh.read(1 << 20), b""):
            h.update(block)
    return {"path": os.path.relpath(LIB_PATH, os.path.dirname(os.path.dirname(_HERE))) if LIB_PATH.startswith(os.path.dirname(os.path.dirname(_HERE))) else LIB_PATH,
            "sha256": h.hexdigest(), "version": lib().gtp_version().decode(), "in_tree": os.path.realpath(LIB_PATH) == os.path.realpath(DEFAULT_LIB_PATH)}


def last_error():
    return lib().gtp_last_error().decode()


def check(rc):
    """Translate a libgtp_b200 return code into the exception the reference raises."""
    if rc == 0:
        return
    msg = last_error()
    if rc == GTP_E_WIDTH_5KM:
        # _modis_interpolator.pyx:33-36
        raise NotImplementedError(msg)
    if rc > 0 and rc != GTP_E_NO_DEVICE:
        raise ValueError(msg)
    raise GtpError(f"libgtp_b200 rc={rc}: {msg}")


def suffix(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32"
    if dtype == np.float64:
        return "f64"
    # what Cython's fused-type dispatch raises for any other buffer dtype
    raise TypeError("No matching signature found")


# -- page-locked host buffers ------------------------------------------------------
class _PinnedPool:
    """Free-list of page-locked blocks keyed by size.

    ``cudaHostAlloc`` costs about as much as copying the buffer, so blocks handed out as
    numpy outputs go back to the pool when the arrays are garbage collected and are
    reused by the next call of the same shape.  Bounded by ``max_cached_bytes``.
    """

    def __init__(self, max_cached_bytes=None):
        if max_cached_bytes is None:
            # GTP_B200_PINNED_CACHE_MB bounds what stays cached after the arrays are gone (default 8 GiB: the
            # lon + lat of two float64 granules at 250 m; geotiepoints_b200.trim() empties it)
            max_cached_bytes = int(float(os.environ.get("GTP_B200_PINNED_CACHE_MB", "8192")) * (1 << 20))
        self.max_cached_bytes = max_cached_bytes
        self._free = {}
        self._order = []               # (nbytes, ptr) of the cached blocks, oldest first
        self._cached = 0
        self._lock = threading.Lock()
        self.pageable_fallbacks = 0

    def _release(self, ptr, nbytes):
        evicted = []
        with self._lock:
            if nbytes > self.max_cached_bytes:
                evicted.append(ptr)
            else:
                # make room by dropping the blocks that have been idle longest (sizes nobody asks for any more)
                while self._cached + nbytes > self.max_cached_bytes and self._order:
                    old_bytes, old_ptr = self._order.pop(0)
                    self._free[old_bytes].remove(old_ptr)
                    self._cached -= old_bytes
                    evicted.append(old_ptr)
                self._free.setdefault(nbytes, []).append(ptr)
                self._order.append((nbytes, ptr))
                self._cached += nbytes
        for p in evicted:
            lib().gtp_host_free(p)

    def empty(self, shape, dtype):
        dtype = np.dtype(dtype)
        nbytes = max(int(np.prod(shape)) * dtype.itemsize, 1)
        ptr = None
        with self._lock:
            blocks = self._free.get(nbytes)
            if blocks:
                ptr = blocks.pop()
                self._order.remove((nbytes, ptr))
                self._cached -= nbytes
        if ptr is None:
            out = _vp()
            rc = lib().gtp_host_alloc(ctypes.byref(out), nbytes)
            if rc != 0:
                # no lockable memory left (ulimit -l, many results alive): give the cache back and try once
                # more, then fall back to pageable memory - the copy is slower, the result the same
                self.trim()
                rc = lib().gtp_host_alloc(ctypes.byref(out), nbytes)
            if rc != 0:
                self.pageable_fallbacks += 1
                return np.empty(shape, dtype)
            ptr = out.value
        buf = (ctypes.c_char * nbytes).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        weakref.finalize(buf, self._release, ptr, nbytes)
        return arr

    def trim(self):
        with self._lock:
            blocks, self._free, self._cached, self._order = self._free, {}, 0, []
        for ptrs in blocks.values():
            for ptr in ptrs:
                lib().gtp_host_free(ptr)


pinned_pool = _PinnedPool()


def trim():
    """Give back what the library and the pinned pool keep between calls: the idle stream / scratch sets of
    the host entry points (~300 MB of device memory each) and the cached page-locked blocks.  For long-lived
    workers that share the GPU or the host's lockable memory with other frameworks."""
    pinned_pool.trim()
    check(lib().gtp_trim())


def set_d2h_relay(device):
    """Route the device-to-host copies of the numpy paths of this process through GPU ``device`` (NVLink peer copy
    + D2H from there); ``None`` or -1 switches it off.  See ``gtp_set_d2h_relay`` in include/gtp_b200.h."""
    check(lib().gtp_set_d2h_relay(-1 if device is None else int(device)))


def device_count():
    return lib().gtp_device_count()


def launch_count():
    return int(lib().gtp_launch_count())


def bind_host_thread_to_gpu(device_index):
    """Pin the calling process to the CPU cores NVML reports as local to ``device_index`` so that
    page-locked staging buffers are first-touched on the NUMA node next to that GPU (a D2H copy
    into the far socket can run several times slower than one into the near socket).
    Returns the CPU set applied, or None when NVML / sched_setaffinity is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, n_words)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None
