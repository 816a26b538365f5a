#!/usr/bin/env python
"""Benchmark of the MODIS geolocation interpolation hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repository (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path

Metric (BASELINE.json): interpolated lon/lat output pixels per second, 1 km -> 250 m, and the
fraction of the HBM roofline.  A *step* is one pass of the CVIIRS interpolator over one batch
of synthetic granules that is already resident in HBM: ``--granules-per-gpu`` (default 36)
MODIS-shaped granules of 2030 x 1354 float64 tie points per GPU -> 36 x 8120 x 5416 output
pixels x (lon, lat).  The per-GPU batch is fixed (weak scaling); at 8 GPUs the job is exactly
the "full day" of 288 granules of BASELINE configs[3], scan-sharded with no collective on
the data path (scans are independent).  One step = one kernel launch per rank.

Prints ONE JSON line (rank 0).  `value` is whole-job px/s with inputs resident in HBM;
`e2e` is the same metric through the public numpy API (pinned host buffers, H2D + D2H inside
the timed region); `roofline` is the dominant kernel against MEASURED_PEAKS.json;
`cpu_baseline` is the reference (oracle/_ref) timed on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "python-geotiepoints_b200"))

import numpy as np  # noqa: E402

METRIC = "interpolated lon/lat pixels/sec (1km->250m)"
UNIT = "px/s"
SCANS_PER_GRANULE = 203
FALLBACK_HBM_GBS = 6650.0            # /opt/skills/guides/B200_PROFILING.md fallback

# The headline workload (default) and the secondary configurations of SURVEY.md 8(d).  The
# driver only runs the default; the others are for `python bench.py --config ...`.
CONFIGS = {
    "cviirs_1km_250m": dict(kind="cviirs", coarse=1000, fine=250, width=0, n_in=3, tie_rows=2030, tie_cols=1354, f=4, out_cols=5416),
    "cviirs_1km_500m": dict(kind="cviirs", coarse=1000, fine=500, width=0, n_in=3, tie_rows=2030, tie_cols=1354, f=2, out_cols=2708),
    "cviirs_5km_1km": dict(kind="cviirs", coarse=5000, fine=1000, width=271, n_in=3, tie_rows=406, tie_cols=271, f=5, out_cols=1354),
    "cviirs_5km_1km_270": dict(kind="cviirs", coarse=5000, fine=1000, width=270, n_in=3, tie_rows=406, tie_cols=270, f=5, out_cols=1354),
    "simple_1km_250m": dict(kind="simple", coarse=1000, fine=250, width=0, n_in=2, tie_rows=2030, tie_cols=1354, f=4, out_cols=5416),
    "simple_1km_500m": dict(kind="simple", coarse=1000, fine=500, width=0, n_in=2, tie_rows=2030, tie_cols=1354, f=2, out_cols=2708),
    # extension: float32 lon/lat + int16 zenith in, float64 math, float32 out (--dtype is ignored)
    "mod03_1km_250m": dict(kind="mixed", coarse=1000, fine=250, width=0, n_in=3, tie_rows=2030, tie_cols=1354, f=4, out_cols=5416),
    # extension: ECEF x / y / z instead of lon / lat (no back-transform, three outputs)
    "ecef_1km_250m": dict(kind="cviirs", ecef=True, coarse=1000, fine=250, width=0, n_in=3, tie_rows=2030, tie_cols=1354, f=4, out_cols=5416),
    # the legacy spline API (geotiepoints.modis1kmto250m / modis5kmto1km), float64, lon + lat only
    "legacy_1km_250m": dict(kind="legacy", code=2, coarse=1000, fine=250, width=0, n_in=2, tie_rows=2030, tie_cols=1354, f=4, out_cols=5416),
    "legacy_5km_1km": dict(kind="legacy", code=0, coarse=5000, fine=1000, width=271, n_in=2, tie_rows=406, tie_cols=271, f=5, out_cols=1354),
    # VII tie_points_geo_interpolation on Cartesian coordinates; the 1 km lon / lat of the synthetic granules
    # serve as tie points (10 tie rows per scan, factor 4): (2030, 1354) -> (203 * 9 * 4, 1353 * 4) per granule
    "vii_geo_cartesian": dict(kind="vii", coarse=1000, fine=250, width=0, n_in=2, tie_rows=2030, tie_cols=1354, f=4, out_cols=5412,
                              out_rows=203 * 36),
    # SURVEY 8f-4, third item: GeoGridInterpolator (default method "linear", extrapolating) with the 1 km lon / lat of the
    # synthetic granules as tie points at fine rows 4 r + 1.5 and fine columns 4 c: (2030, 1354) -> (8120, 5416) per granule
    "geogrid_linear": dict(kind="grid", coarse=1000, fine=250, width=0, n_in=2, tie_rows=2030, tie_cols=1354, f=4, out_cols=5416),
    "mod03_ecef_1km_250m": dict(kind="mixed", ecef=True, coarse=1000, fine=250, width=0, n_in=3, tie_rows=2030, tie_cols=1354, f=4, out_cols=5416),
}
TIE_ROWS, TIE_COLS = 2030, 1354
OUT_ROWS, OUT_COLS = 8120, 5416
PX_PER_GRANULE = OUT_ROWS * OUT_COLS


def px_per_granule(cfg):
    return cfg.get("out_rows", cfg["tie_rows"] * cfg["f"]) * cfg["out_cols"]


def algorithmic_bytes_per_px(itemsize, cfg=None):
    """Each input element read once, each output written once: 17.5 B/px for
    CVIIRS 1 km -> 250 m float64."""
    cfg = cfg or CONFIGS["cviirs_1km_250m"]
    n_out = 3 if cfg.get("ecef") else 2
    if cfg["kind"] == "mixed":      # float32 out; float32 + float32 + int16 in
        return n_out * 4 + (4 + 4 + 2) * (cfg["tie_rows"] * cfg["tie_cols"]) / px_per_granule(cfg)
    return n_out * itemsize + cfg["n_in"] * itemsize * (cfg["tie_rows"] * cfg["tie_cols"]) / px_per_granule(cfg)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--dtype", choices=["f64", "f32"], default="f64")
    ap.add_argument("--granules-per-gpu", type=int, default=36)
    ap.add_argument("--e2e-granules", type=int, default=4)
    ap.add_argument("--config", choices=sorted(CONFIGS), default="cviirs_1km_250m",
                    help="workload; the default is the headline metric, the others are SURVEY 8(d) secondary configs")
    ap.add_argument("--kernel", choices=["auto", "generic"], default="auto",
                    help="auto = what the public API runs; generic = force the rounding-faithful kernel")
    ap.add_argument("--sustained-seconds", type=float, default=1.0,
                    help="extra load before the `sustained` measurement (0 = skip it)")
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="do not pin each rank to the CPU cores local to its GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-relay", action="store_true", help="N > 1: do not try the NVLink relay of the device-to-host copies")
    ap.add_argument("--no-secondary", action="store_true", help="skip the secondary configurations of the default run")
    ap.add_argument("--secondary-granules", type=int, default=16)
    return ap.parse_args()


# ------------------------------------------------------------------------------------
# synthetic inputs
# ------------------------------------------------------------------------------------
def make_granules(first, count, dtype, cfg=None):
    """`count` consecutive synthetic granules of the day, stacked along the scan axis."""
    from concurrent.futures import ThreadPoolExecutor
    from geotiepoints_b200 import testing
    cfg = cfg or CONFIGS["cviirs_1km_250m"]

    def one(g):
        arrs = testing.modis_granule(g % 288, dtype=dtype)
        if cfg["coarse"] == 5000:
            arrs = testing.subsample_5km(*arrs, width=cfg["width"])
        return arrs
    with ThreadPoolExecutor(min(8, max(1, count))) as pool:
        parts = list(pool.map(one, range(first, first + count)))
    return tuple(np.concatenate([p[k] for p in parts]) for k in range(cfg["n_in"]))


# ------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons sampled every few ms through NVML (nvidia-ml-py) in a
    background thread, so that even a 50 ms timed region gets its own samples; falls back to
    `nvidia-smi -lms` when NVML is not importable."""
    REASONS = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
               "hw_power_brake_slowdown": 0x80}

    def __init__(self, gpu_index, period_s=0.004):
        self.gpu_index, self.period_s = gpu_index, period_s
        self.samples = []          # (t, sm_mhz, power_w, reasons_bitmask)
        self.sm_max = None
        self._stop = threading.Event()
        self._thread = None
        self.source = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            handle = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"

            def loop():
                while not self._stop.is_set():
                    try:
                        mhz = pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
                        watts = pynvml.nvmlDeviceGetPowerUsage(handle) / 1000.0
                        try:
                            mask = pynvml.nvmlDeviceGetCurrentClocksEventReasons(handle)
                        except AttributeError:
                            mask = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                        self.samples.append((time.perf_counter(), float(mhz), watts, int(mask)))
                    except Exception:
                        pass
                    self._stop.wait(self.period_s)
            self._thread = threading.Thread(target=loop, daemon=True)
            self._thread.start()
        except Exception:
            self._start_smi()

    def _start_smi(self):
        self.source = "nvidia-smi"
        query = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active"
        try:
            proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={query}",
                                     "--format=csv,noheader,nounits", "-lms", "50"],
                                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.source = None
            return

        def loop():
            for line in proc.stdout:
                parts = [x.strip() for x in line.split(",")]
                try:
                    self.sm_max = float(parts[1])
                    self.samples.append((time.perf_counter(), float(parts[0]), float(parts[2]), int(parts[3], 16)))
                except (ValueError, IndexError):
                    continue
                if self._stop.is_set():
                    break
            proc.terminate()
        self._thread = threading.Thread(target=loop, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def window(self, t0, t1):
        """Summary of the samples taken in [t0, t1] (perf_counter seconds)."""
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML / nvidia-smi"], "samples": 0}
        rows = [r for r in self.samples if t0 <= r[0] <= t1]
        if not rows:       # region shorter than the sampling period: take the nearest sample
            rows = sorted(self.samples, key=lambda r: abs(r[0] - 0.5 * (t0 + t1)))[:1]
        mask = 0
        for r in rows:
            mask |= r[3]
        return {"sm_mhz": float(np.median([r[1] for r in rows])) if rows else None,
                "sm_max_mhz": self.sm_max, "power_w": float(np.median([r[2] for r in rows])) if rows else None,
                "samples": len(rows), "source": self.source,
                "reasons": sorted(k for k, bit in self.REASONS.items() if mask & bit)}


# ------------------------------------------------------------------------------------
# CPU baseline = the reference itself on the host cores
# ------------------------------------------------------------------------------------
def load_reference():
    """(kind, callable(lon, lat, satz) -> (lon, lat)) for float64 1 km -> 250 m.

    kind == "reference": the reference's Cython code built from /root/reference into
    oracle/_ref by oracle/build_ref.py - the float64 twin with the 2-line dtype patch the
    stock CVIIRS path needs to accept float64 at all.  Falls back to the C
    restatement (kind == "port") when oracle/_ref is not importable.
    """
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    try:
        if ref_dir not in sys.path:
            sys.path.insert(0, ref_dir)
        import geotiepoints._modis_interpolator as stock
        import geotiepoints_f64._modis_interpolator as f64

        def run(lon, lat, satz):
            mod = stock if lon.dtype == np.float32 else f64
            return mod.interpolate(lon, lat, satz, coarse_resolution=1000, fine_resolution=250)
        return "reference", run
    except ImportError:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import gtp_oracle
        return "port", lambda lon, lat, satz: gtp_oracle.cviirs(lon, lat, satz, 1000, 250)


def time_reference_step(run, lon, lat, satz, threads):
    """One pass of the reference over one granule, scans split over `threads` host threads
    (the Cython scan loop is nogil: _modis_interpolator.pyx:175; stands in for dask map_blocks)."""
    from concurrent.futures import ThreadPoolExecutor
    n_scans = lon.shape[0] // 10
    bounds = [(k * n_scans // threads, (k + 1) * n_scans // threads) for k in range(threads)]
    bounds = [(a, b) for a, b in bounds if b > a]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(len(bounds)) as pool:
        list(pool.map(lambda ab: run(lon[ab[0] * 10:ab[1] * 10], lat[ab[0] * 10:ab[1] * 10], satz[ab[0] * 10:ab[1] * 10]), bounds))
    return time.perf_counter() - t0


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return None


def cpu_baseline(dtype, reps=3):
    """The reference on this box's host cores: all threads on one granule (the reported value) and one
    thread on 40 scans of it (SURVEY.md 8d asks for both)."""
    kind, run = load_reference()
    lon, lat, satz = make_granules(0, 1, dtype)
    threads = host_threads()
    time_reference_step(run, lon[:200], lat[:200], satz[:200], min(threads, 4))      # warm-up
    best = min(time_reference_step(run, lon, lat, satz, threads) for _ in range(reps))
    rows1 = 400
    best1 = min(time_reference_step(run, lon[:rows1], lat[:rows1], satz[:rows1], 1) for _ in range(reps))
    return {"value": PX_PER_GRANULE / best, "unit": UNIT, "cores": threads, "kind": kind,
            "single_thread_value": (rows1 * 4 * OUT_COLS) / best1, "cpu_model": cpu_model(),
            "sample": f"1 synthetic granule 2030x1354 {np.dtype(dtype).name} 1km->250m, best of {reps}, "
                      f"scans split over {threads} threads" + (" (f64 = reference + 2-line dtype patch)" if kind == "reference" and dtype == np.float64 else "")
                      + f"; single_thread_value: {rows1 // 10} scans on 1 thread"}


def run_reference_arm(args, dtype):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, run = load_reference()
    lon, lat, satz = make_granules(0, 1, dtype)
    threads = host_threads()
    for _ in range(max(1, min(args.warmup, 2))):
        time_reference_step(run, lon, lat, satz, threads)
    t = [time_reference_step(run, lon, lat, satz, threads) for _ in range(args.steps)]
    total = sum(t)
    value = PX_PER_GRANULE * args.steps / total
    sample = (f"each step = 1 synthetic granule 2030x1354 {np.dtype(dtype).name} 1km->250m (bounded sample of the "
              f"workload), scans split over {threads} host threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": "CVIIRS 1km->250m, synthetic MODIS granules 2030x1354 tie points, " + args.dtype,
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------
# this repository's arm
# ------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------
# secondary configurations (SURVEY.md 8d / BASELINE configs[2] and [4]) inside the default run
# ------------------------------------------------------------------------------------
SECONDARY = [("cviirs_1km_250m", "f32"), ("cviirs_1km_500m", "f64"), ("cviirs_1km_500m", "f32"),
             ("cviirs_5km_1km", "f64"), ("cviirs_5km_1km", "f32"), ("simple_1km_250m", "f64"), ("simple_1km_250m", "f32"),
             ("simple_1km_500m", "f64"), ("mod03_1km_250m", "f32"), ("legacy_1km_250m", "f64"), ("legacy_5km_1km", "f64"),
             ("vii_geo_cartesian", "f64"), ("geogrid_linear", "f64")]


def measure_secondary(name, dtype_name, base_in, granules, device, local_rank, peak, steps=5, warmup=3, repeat=1):
    """One secondary configuration with the data resident in HBM: `steps` timed launches (CUDA events on the
    launch stream) over `granules` synthetic granules (working set >> L2), roofline fraction from the
    algorithmic bytes, parity of 6 scans spread over the batch against the oracle.
    base_in: float64 (lon, lat, satz) of `granules` 1 km granules."""
    import ctypes
    import torch
    from geotiepoints_b200 import _capi, testing
    cfg = CONFIGS[name]
    mixed = cfg["kind"] == "mixed"
    dtype = np.float32 if (dtype_name == "f32" or mixed) else np.float64
    arrs = base_in
    if cfg["kind"] == "legacy" and cfg["coarse"] == 5000:
        arrs = [a[2::5, 2::5] for a in arrs]                   # geotiepoints/tests: 5 km tie points of a 1 km granule
    elif cfg["coarse"] == 5000:
        arrs = testing.subsample_5km(*arrs, width=cfg["width"])
    arrs = [np.ascontiguousarray(a.astype(dtype)) for a in arrs[:cfg["n_in"]]]
    if mixed:
        arrs[2] = np.round(base_in[2] / 0.01).astype(np.int16)
    if repeat > 1:                                              # scans are independent: the batch `repeat` times over
        arrs = [np.tile(a, (repeat, 1)) for a in arrs]
        granules *= repeat
    rows_per_scan = 10 if cfg["coarse"] == 1000 else 2
    n_scans = arrs[0].shape[0] // rows_per_scan
    out_rows_per_scan = (rows_per_scan - 1) * cfg["f"] if cfg["kind"] == "vii" else rows_per_scan * cfg["f"]
    tdtype = torch.float64 if dtype == np.float64 else torch.float32
    d_in = [torch.from_numpy(a).to(device) for a in arrs]
    outs = [torch.empty((n_scans * out_rows_per_scan, cfg["out_cols"]), dtype=tdtype, device=device) for _ in range(2)]
    lib = _capi.lib()
    stream = torch.cuda.current_stream(device)
    st = ctypes.c_void_p(stream.cuda_stream)
    ptrs = [t.data_ptr() for t in d_in]
    optrs = [o.data_ptr() for o in outs]
    if mixed:
        def step():
            _capi.check(lib.gtp_cviirs_interp_mixed(*ptrs, 1, 0.01, 0.0, n_scans, cfg["coarse"], cfg["fine"], 0, *optrs, 0, local_rank, st))
    elif cfg["kind"] == "legacy":
        def step():
            _capi.check(lib.gtp_legacy_interp_f64(*ptrs, n_scans, cfg["code"], *optrs, local_rank, st))
    elif cfg["kind"] == "vii":
        def step():
            _capi.check(lib.gtp_vii_geo_interp_f64(*ptrs, n_scans * 10, cfg["tie_cols"], 10, cfg["f"], 1, 0.8, *optrs, local_rank, st))
    elif cfg["kind"] == "grid":
        n_tie = n_scans * 10
        grid_axes = [4.0 * np.arange(n_tie) + 1.5, 4.0 * np.arange(cfg["tie_cols"]), np.arange(4.0 * n_tie), np.arange(4.0 * cfg["tie_cols"])]
        d_axes = [torch.from_numpy(a).to(device) for a in grid_axes]
        ty, tx, fy, fx = [t.data_ptr() for t in d_axes]

        def step():
            _capi.check(lib.gtp_geogrid_interp_f64(ty, n_tie, tx, cfg["tie_cols"], *ptrs, fy, 4 * n_tie, fx, 4 * cfg["tie_cols"],
                                                   1, 1, 0, float("nan"), *optrs, local_rank, st))
    elif cfg["kind"] == "cviirs":
        fn = getattr(lib, "gtp_cviirs_interp_" + dtype_name)

        def step():
            _capi.check(fn(*ptrs, n_scans, cfg["coarse"], cfg["fine"], cfg["width"], *optrs, local_rank, st))
    else:
        fn = getattr(lib, "gtp_simple_interp_" + dtype_name)

        def step():
            _capi.check(fn(*ptrs, n_scans, cfg["tie_cols"], cfg["coarse"] // cfg["fine"], *optrs, local_rank, st))
    for _ in range(warmup):
        step()
    torch.cuda.synchronize(device)
    before = _capi.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        a.record(stream)
        step()
        b.record(stream)
    torch.cuda.synchronize(device)
    launches = _capi.launch_count() - before
    ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    px = n_scans * out_rows_per_scan * cfg["out_cols"]
    bpp = algorithmic_bytes_per_px(np.dtype(dtype).itemsize, cfg)
    res = {"config": name, "dtype": "f32+i16->f32 (f64 math)" if mixed else dtype_name, "value": px / (ms * 1e-3), "unit": UNIT,
           "kernel_ms": ms, "steps": steps, "granules": granules, "gpu_launches": int(launches),
           "algorithmic_bytes_per_px": bpp, "roofline_frac": bpp * px / (ms * 1e-3) / 1e9 / peak,
           "working_set_gb": bpp * px / 1e9}
    try:
        import gtp_oracle
        from parity import compare
        picks = np.unique(np.linspace(0, n_scans - 1, 6).astype(np.int64))
        in_rows = np.concatenate([np.arange(s * rows_per_scan, (s + 1) * rows_per_scan) for s in picks])
        out_rows = np.concatenate([np.arange(s * out_rows_per_scan, (s + 1) * out_rows_per_scan) for s in picks])
        sub = [np.ascontiguousarray(a[in_rows]) for a in arrs]
        idx = torch.from_numpy(out_rows).to(device)
        with np.errstate(all="ignore"):
            if cfg["kind"] == "grid":                          # the fine rows of a "scan" also see the tie rows next to it
                import grid_oracle
                parts = []
                for s_ in picks:
                    r0, r1 = max(10 * s_ - 1, 0), min(10 * s_ + 11, n_tie)
                    parts.append(grid_oracle.geogrid_interpolate(grid_axes[0][r0:r1], grid_axes[1], arrs[0][r0:r1], arrs[1][r0:r1],
                                                                 grid_axes[2][40 * s_:40 * s_ + 40], grid_axes[3]))
                exp = tuple(np.concatenate([p_[k] for p_ in parts]) for k in (0, 1))
            elif mixed:
                exp = gtp_oracle.cviirs(sub[0].astype(np.float64), sub[1].astype(np.float64), sub[2].astype(np.float64) * 0.01,
                                        1000, cfg["fine"], 0, n_threads=8)
                exp = tuple(a.astype(np.float32) for a in exp)
            elif cfg["kind"] == "vii":
                import vii_oracle
                xyz = vii_oracle.tie_points_interpolation(list(vii_oracle.lonlat2xyz(*sub)), 10, cfg["f"])
                exp = vii_oracle.xyz2lonlat(*xyz, 0.8)
            elif cfg["kind"] == "legacy":
                import legacy_oracle
                exp = legacy_oracle.interpolate(*sub, {0: "5km_1km", 1: "1km_500m", 2: "1km_250m"}[cfg["code"]])
            elif cfg["kind"] == "cviirs":
                exp = gtp_oracle.cviirs(*sub, cfg["coarse"], cfg["fine"], cfg["width"], n_threads=8)
            else:
                exp = gtp_oracle.simple(*sub, cfg["coarse"], cfg["fine"], n_threads=8)
        par = compare(outs[0][idx].cpu().numpy(), outs[1][idx].cpu().numpy(), *exp)
        keep = ("n", "nan_pattern_equal", "max_dlat", "max_dlon_outside_band", "max_dlon_in_band", "n_in_band",
                "n_in_band_over_conditioning_bound", "bit_exact_fraction", "n_over_tol_and_ulp")
        res["parity"] = {k: par[k] for k in keep if k in par}
    except Exception as exc:
        res["parity"] = {"error": repr(exc)}
    del d_in, outs
    torch.cuda.empty_cache()
    return res


class _LazyBlocks:
    """Stand-in for ``dask.array`` in `e2e.dask` when dask is not installed (it is not in this image): whole-array
    inputs with ``chunks``, ``rechunk`` and ``map_blocks(new_axis=[0])``, evaluated block by block on
    ``compute()`` and assembled with ``np.concatenate`` as dask does.  The block results are cached so that
    computing lons and lats evaluates every block once (what ``dask.compute(lons, lats)`` does)."""
    def __init__(self, thunk, shape, chunks, dtype):
        self._thunk, self.shape, self.dtype, self.ndim = thunk, tuple(shape), np.dtype(dtype), len(shape)
        self.chunks = tuple(tuple(c) if isinstance(c, (tuple, list)) else ((n,) if c == -1 else
                            (c,) * (n // c) + ((n % c,) if n % c else ())) for c, n in zip(chunks, shape))
        self._cache = None

    def rechunk(self, chunks):
        return _LazyBlocks(self._thunk, self.shape, chunks, self.dtype)

    def compute(self):
        if self._cache is None:
            self._cache = self._thunk(self)
        return self._cache

    def __getitem__(self, k):
        parent = self
        return _LazyBlocks(lambda _s: parent.compute()[k], self.shape[1:], self.chunks[1:], self.dtype)

    @staticmethod
    def from_array(a, chunks):
        return _LazyBlocks(lambda _s: a, a.shape, chunks, a.dtype)

    @staticmethod
    def map_blocks(func, *arrays, new_axis=None, chunks=None, dtype=None, meta=None, **kwargs):
        first = arrays[0]
        chunks = tuple((c,) if isinstance(c, int) else tuple(c) for c in chunks)

        def thunk(_s):
            data = [a.compute() for a in arrays]
            pieces, r0 = [], 0
            for rc in first.chunks[0]:
                pieces.append(func(*[d[r0:r0 + rc] for d in data], **kwargs))
                r0 += rc
            return pieces[0] if len(pieces) == 1 else np.concatenate(pieces, axis=1)
        return _LazyBlocks(thunk, tuple(sum(c) for c in chunks), chunks, dtype)


def e2e_lazy(h_in, steps, chunk_rows, rows_per_array=None):
    """1 km -> 250 m through the dask-shaped call: lazy inputs -> scanline_mapblocks -> compute().
    rows_per_array: cut the inputs into arrays of that many tie rows (one granule = 2030) that are interpolated one
    after the other, the way a reader hands them over; None = one array of all rows."""
    from geotiepoints_b200 import _modis_utils, modisinterpolator
    try:
        import dask
        import dask.array as da

        def lazy(a):
            return da.from_array(a, chunks=(chunk_rows, -1))

        def compute(lons, lats):
            return dask.compute(lons, lats, scheduler="threads")
        kind = "dask " + dask.__version__
    except ImportError:
        import types
        saved = _modis_utils.da
        _modis_utils.da = types.SimpleNamespace(map_blocks=_LazyBlocks.map_blocks, Array=_LazyBlocks)

        def lazy(a):
            return _LazyBlocks.from_array(a, (chunk_rows, -1))

        def compute(lons, lats):
            return lons.compute(), lats.compute()
        kind = "stand-in block scheduler (dask is not installed in this image)"
    rows = h_in[0].shape[0]
    step_rows = rows_per_array or rows
    pieces = [[a[r0:r0 + step_rows] for a in h_in] for r0 in range(0, rows, step_rows)]

    def one_pass():
        checksum = 0.0
        for piece in pieces:
            lons, lats = modisinterpolator.modis_1km_to_250m(*[lazy(a) for a in piece])
            res = compute(lons, lats)
            checksum = float(res[0][-1, -1]) + float(res[1][-1, -1])
            del res, lons, lats
        return checksum
    try:
        one_pass()
        t0 = time.perf_counter()
        for _ in range(steps):
            checksum = one_pass()
        dt = time.perf_counter() - t0
    finally:
        if "saved" in locals():
            _modis_utils.da = saved
    return dt, checksum, kind


def run_b200_arm(args, dtype):
    import ctypes
    import torch
    import torch.distributed as dist
    from geotiepoints_b200 import _capi, modisinterpolator, sharding

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world:
        raise SystemExit(f"--gpus {args.gpus} needs {args.gpus} ranks: launch with python -m torch.distributed.run "
                         f"--nnodes=1 --nproc-per-node {args.gpus} --master-addr 127.0.0.1 bench.py --gpus {args.gpus} ...")
    if _capi.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: geotiepoints_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    bound_cpus = None if args.no_numa_bind else _capi.bind_host_thread_to_gpu(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    cfg = CONFIGS[args.config]
    if cfg["kind"] in ("legacy", "vii", "grid"):
        raise SystemExit("the legacy_* / vii_* / geogrid_* configurations are measured inside the default run (`secondary`), not as --config")
    headline = args.config == "cviirs_1km_250m"
    mixed = cfg["kind"] == "mixed"
    ecef = bool(cfg.get("ecef"))
    n_out = 3 if ecef else 2
    if mixed:
        dtype, args.dtype = np.float32, "f32"
    tdtype = torch.float64 if dtype == np.float64 else torch.float32
    itemsize = np.dtype(dtype).itemsize
    G = args.granules_per_gpu
    # rank r owns the contiguous scan range of granules [r*G, (r+1)*G) of the day
    g_lo, g_hi = sharding.granule_range(rank, world, G * world)
    host_in = make_granules(g_lo, g_hi - g_lo, dtype, cfg)
    if mixed:       # MOD03 layout: SensorZenith as int16 * 0.01
        host_in = (host_in[0], host_in[1], np.round(host_in[2].astype(np.float64) / 0.01).astype(np.int16))
    lon, lat = host_in[0], host_in[1]
    rows_per_scan = 10 if cfg["coarse"] == 1000 else 2
    n_scans = lon.shape[0] // rows_per_scan
    out_rows_per_scan = rows_per_scan * cfg["f"]
    d_in = [torch.from_numpy(a).to(device) for a in host_in]
    outs = [torch.empty((n_scans * out_rows_per_scan, cfg["out_cols"]), dtype=tdtype, device=device) for _ in range(n_out)]
    out_lon, out_lat = outs[0], outs[1]
    out_ptrs = [o.data_ptr() for o in outs]
    lib = _capi.lib()
    suffix = "f64" if dtype == np.float64 else "f32"
    generic = args.kernel == "generic" and (suffix == "f64" or (cfg["kind"] == "cviirs" and not ecef)) and not mixed
    xyz = "xyz_" if ecef else ""
    fn = getattr(lib, f"gtp_cviirs_interp_{xyz}mixed") if mixed else \
        getattr(lib, f"gtp_{cfg['kind']}_interp_{xyz}" + (("generic_" + suffix) if generic else suffix))
    stream = torch.cuda.current_stream(device)
    if cfg["kind"] == "cviirs":
        scalars = (cfg["coarse"], cfg["fine"], cfg["width"])
    else:
        scalars = (cfg["tie_cols"], cfg["coarse"] // cfg["fine"])

    def step():
        if mixed:       # the ECEF form has no coarse_width argument
            rc = fn(*[t.data_ptr() for t in d_in], 1, 0.01, 0.0, n_scans, cfg["coarse"], cfg["fine"], *(() if ecef else (0,)),
                    *out_ptrs, 0, local_rank, ctypes.c_void_p(stream.cuda_stream))
        else:
            rc = fn(*[t.data_ptr() for t in d_in], n_scans, *scalars,
                    *out_ptrs, local_rank, ctypes.c_void_p(stream.cuda_stream))
        _capi.check(rc)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # -- timed region: exactly K steps, one kernel launch each -------------------------
    launches_before = _capi.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.perf_counter()
    t_begin.record(stream)
    for a, b in ev:
        a.record(stream)
        step()
        b.record(stream)
    t_end.record(stream)
    barrier()
    wall1 = time.perf_counter()
    elapsed_ms = t_begin.elapsed_time(t_end)
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    launches = _capi.launch_count() - launches_before
    clocks = sampler.window(wall0, wall1) if rank == 0 else None

    # -- sustained regime (reported beside the contract number, not instead of it): the kernel is
    # FP64-heavy and a B200 reaches its 1 kW power cap after a few hundred ms of it ------------
    sustained = None
    if args.sustained_seconds > 0:
        t_s = time.perf_counter()
        while time.perf_counter() - t_s < args.sustained_seconds:
            for _ in range(8):
                step()
            torch.cuda.synchronize(device)
        n_sus = max(8, int(0.4 / max(kernel_ms * 1e-3, 1e-4)))
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        wall2 = time.perf_counter()
        s0.record(stream)
        for _ in range(n_sus):
            step()
        s1.record(stream)
        torch.cuda.synchronize(device)
        wall3 = time.perf_counter()
        sus_ms = sharding.max_over_ranks(s0.elapsed_time(s1) / n_sus, dist if world > 1 else None)
        sustained = {"ms_per_step": sus_ms, "steps": n_sus, "after_seconds_of_load": args.sustained_seconds,
                     "clocks": sampler.window(wall2, wall3) if rank == 0 else None}
    if rank == 0:
        sampler.stop()
    elapsed_ms = sharding.max_over_ranks(elapsed_ms, dist if world > 1 else None)
    kernel_ms_max = sharding.max_over_ranks(kernel_ms, dist if world > 1 else None)

    px_per_rank_step = n_scans * out_rows_per_scan * cfg["out_cols"]
    total_px = px_per_rank_step * world * args.steps
    value = total_px / (elapsed_ms * 1e-3)

    # -- parity spot check (not timed): 12 scans spread over this rank's batch (every latitude band it
    # holds, incl. polar and antimeridian scans) against the oracle --------------------------------
    parity = None
    if rank == 0:
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
            import gtp_oracle
            from parity import compare
            picks = np.unique(np.linspace(0, n_scans - 1, 12).astype(np.int64))
            in_rows = np.concatenate([np.arange(s * rows_per_scan, (s + 1) * rows_per_scan) for s in picks])
            out_rows = np.concatenate([np.arange(s * out_rows_per_scan, (s + 1) * out_rows_per_scan) for s in picks])
            sub = [np.ascontiguousarray(a[in_rows]) for a in host_in]
            idx = torch.from_numpy(out_rows).to(device)
            with np.errstate(all="ignore"):
                if ecef:
                    from parity import compare_xyz
                    if mixed:
                        exp = gtp_oracle.cviirs_xyz(sub[0].astype(np.float64), sub[1].astype(np.float64),
                                                    sub[2].astype(np.float64) * 0.01, 1000, cfg["fine"], 0, n_threads=8)
                        exp = tuple(a.astype(np.float32) for a in exp)
                    else:
                        exp = gtp_oracle.cviirs_xyz(*sub, cfg["coarse"], cfg["fine"], cfg["width"], n_threads=8)
                    parity = compare_xyz([o[idx].cpu().numpy() for o in outs], exp)
                elif mixed:       # the float64 reference on the widened inputs, rounded once
                    exp = gtp_oracle.cviirs(sub[0].astype(np.float64), sub[1].astype(np.float64),
                                            sub[2].astype(np.float64) * 0.01, 1000, cfg["fine"], 0, n_threads=8)
                    exp = tuple(a.astype(np.float32) for a in exp)
                elif cfg["kind"] == "cviirs":
                    exp = gtp_oracle.cviirs(*sub, cfg["coarse"], cfg["fine"], cfg["width"], n_threads=8)
                else:
                    exp = gtp_oracle.simple(*sub, cfg["coarse"], cfg["fine"], n_threads=8)
            if not ecef:
                parity = compare(out_lon[idx].cpu().numpy(), out_lat[idx].cpu().numpy(), *exp)
            parity["scans_checked"] = [int(s) for s in picks]
        except Exception as exc:  # the checker must never take the measurement down
            parity = {"error": repr(exc)}

    # -- end to end through the public API: pinned host buffers, H2D + D2H inside --------
    e2e = None
    if not args.no_e2e:
        from geotiepoints_b200 import simple_modis_interpolator
        from geotiepoints_b200._modis_interpolator import interpolate as cviirs_interpolate
        if mixed:
            from geotiepoints_b200 import mod03

            def public_api(*arrs):
                return mod03.modis_1km_to_250m(*arrs, zenith_scale=0.01, ecef=ecef)
        elif ecef:
            from geotiepoints_b200 import ecef as ecef_api

            def public_api(*arrs):
                return ecef_api.interpolate(*arrs, cfg["coarse"], cfg["fine"], cfg["width"])
        elif cfg["kind"] == "cviirs":
            def public_api(*arrs):
                return cviirs_interpolate(*arrs, coarse_resolution=cfg["coarse"], fine_resolution=cfg["fine"],
                                          coarse_scan_width=cfg["width"])
            if headline:
                public_api = modisinterpolator.modis_1km_to_250m
        else:
            def public_api(*arrs):
                return simple_modis_interpolator.interpolate_geolocation_cartesian(
                    *arrs, coarse_resolution=cfg["coarse"], fine_resolution=cfg["fine"])
        ge = min(args.e2e_granules, G)
        rows = ge * cfg["tie_rows"]
        h_in = []
        for a in host_in:
            pinned = _capi.pinned_pool.empty((rows, cfg["tie_cols"]), a.dtype)
            pinned[...] = a[:rows]
            h_in.append(pinned)
        res = public_api(*h_in)      # warm-up (pinned pool, stream pool)
        del res
        e2e_steps = max(1, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            res = public_api(*h_in)
            checksum = float(res[0][-1, -1]) + float(res[1][-1, -1])
            del res
        torch.cuda.synchronize(device)
        dt = time.perf_counter() - t0
        dt = sharding.max_over_ranks(dt, dist if world > 1 else None)
        e2e_px = ge * px_per_granule(cfg)
        # -- NVLink relay of the device-to-host copies (world > 1): PCIe root ports of a multi-GPU host are not
        # equally fast; ranks on slow ports send their results through a GPU on a fast one (gtp_set_d2h_relay)
        relay_info, relay_on = None, False
        if world > 1 and headline and not args.no_relay:
            probe_d = torch.empty(256 << 20, dtype=torch.uint8, device=device)
            probe_h = _capi.pinned_pool.empty((256 << 20,), np.uint8)
            probe_t = torch.from_numpy(probe_h)
            probe_t.copy_(probe_d, non_blocking=True)
            barrier()
            t0 = time.perf_counter()
            for _ in range(4):
                probe_t.copy_(probe_d, non_blocking=True)
            torch.cuda.synchronize(device)
            my_rate = 4 * (256 << 20) / (time.perf_counter() - t0) / 1e9
            rates_t = [torch.zeros(1, dtype=torch.float64, device=device) for _ in range(world)]
            dist.all_gather(rates_t, torch.tensor([my_rate], dtype=torch.float64, device=device))
            rates = [float(r.item()) for r in rates_t]
            del probe_d, probe_t, probe_h
            fast = [r for r in range(world) if rates[r] >= 0.8 * max(rates)]
            slow = [r for r in range(world) if r not in fast]
            relay_of = {r: fast[i % len(fast)] for i, r in enumerate(slow)}
            relay_info = {"d2h_gbs_per_rank_all_copying": [round(r, 1) for r in rates],
                          "relay_of_rank": {str(k): v for k, v in relay_of.items()},
                          "how": "ranks whose concurrent pinned D2H rate is < 0.8 x the best send their results over NVLink to a "
                                 "staging buffer on a fast-port GPU and copy to the host from there (gtp_set_d2h_relay)"
                                 if slow else "all ports within 20 % of each other: no relay"}
            if slow:
                _capi.set_d2h_relay(relay_of.get(rank))
                res = public_api(*h_in)
                del res
                barrier()
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    res = public_api(*h_in)
                    rsum = float(res[0][-1, -1]) + float(res[1][-1, -1])
                    del res
                torch.cuda.synchronize(device)
                rdt = sharding.max_over_ranks(time.perf_counter() - t0, dist if world > 1 else None)
                relay_info.update({"value": e2e_px * world * e2e_steps / rdt, "unit": UNIT, "ms_per_step": 1e3 * rdt / e2e_steps,
                                   "steps": e2e_steps, "checksum_equal": rsum == checksum})
                relay_on = relay_info["checksum_equal"] and rdt < dt
                if not relay_on:
                    _capi.set_d2h_relay(None)
        # the same through the dask-shaped call (lazy inputs, whole-scan blocks, compute), and the float32-out
        # (mod03) and ECEF forms of the call, which move half / 1.5 x the bytes (headline config only)
        extra_e2e = {}
        if headline:
            lazy_steps = max(1, min(e2e_steps, 2))
            # (a) granule by granule, the way a reader (Satpy) hands them over: 2030 rows < the chunk size of the
            # reference's tests (4096) = one whole-scan block per array, nothing to assemble afterwards;
            # (b) all rows as ONE array: two blocks, which the scheduler concatenates into fresh arrays
            ldt, lsum, lkind = e2e_lazy(h_in, lazy_steps, 4096, rows_per_array=TIE_ROWS)
            ldt = sharding.max_over_ranks(ldt, dist if world > 1 else None)
            extra_e2e["dask"] = {"value": e2e_px * world * lazy_steps / ldt, "unit": UNIT, "ms_per_step": 1e3 * ldt / lazy_steps,
                                 "steps": lazy_steps, "scheduler": lkind, "chunks": "4096 x 4096 as in the reference's tests -> one whole-scan block per granule",
                                 "sample": f"modisinterpolator.modis_1km_to_250m on lazy inputs + compute of lons and lats, {ge} granules one "
                                           "after the other", "checksum": lsum}
            cdt, csum_, _ = e2e_lazy(h_in, 1, 4096)
            cdt = sharding.max_over_ranks(cdt, dist if world > 1 else None)
            extra_e2e["dask_one_array"] = {"value": e2e_px * world / cdt, "unit": UNIT, "ms_per_step": 1e3 * cdt, "steps": 1,
                                           "chunks": "4096 rows (rechunked to 4090 = whole scans) x full width: 2 blocks",
                                           "sample": f"the same with the {ge} granules as ONE lazy array; includes the scheduler's assembly of the "
                                                     "blocks into the result arrays (np.concatenate of 2.8 GB)", "checksum": csum_}
            from geotiepoints_b200 import ecef as ecef_api, mod03
            m_in = [_capi.pinned_pool.empty((rows, cfg["tie_cols"]), dt_) for dt_ in (np.float32, np.float32, np.int16)]
            m_in[0][...] = h_in[0]; m_in[1][...] = h_in[1]; m_in[2][...] = np.round(np.asarray(h_in[2], np.float64) / 0.01)
            for key, call, nbytes_out in (
                    ("mod03_f32", lambda: mod03.modis_1km_to_250m(*m_in), 2 * 4),
                    ("ecef", lambda: ecef_api.modis_1km_to_250m(*h_in), 3 * itemsize)):
                res = call()
                del res
                barrier()
                t0 = time.perf_counter()
                for _ in range(lazy_steps):
                    res = call()
                    csum = float(res[0][-1, -1]) + float(res[-1][-1, -1])
                    del res
                torch.cuda.synchronize(device)
                xdt = sharding.max_over_ranks(time.perf_counter() - t0, dist if world > 1 else None)
                extra_e2e[key] = {"value": e2e_px * world * lazy_steps / xdt, "unit": UNIT, "ms_per_step": 1e3 * xdt / lazy_steps,
                                  "steps": lazy_steps, "d2h_bytes_per_step": int(nbytes_out * e2e_px), "checksum": csum}
            del m_in
        e2e = {"value": e2e_px * world * e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(sum(a.dtype.itemsize for a in host_in) * rows * cfg["tie_cols"]),
               "d2h_bytes_per_step": int(n_out * e2e_px * itemsize),
               "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps,
               "host_cpus_bound_to_gpu": len(bound_cpus) if bound_cpus else None,
               "sample": f"{ge} granules per GPU per call through the public numpy API "
                         f"({'geotiepoints_b200.modisinterpolator.modis_1km_to_250m' if headline else args.config}), "
                         f"pinned host inputs, pinned host outputs",
               "checksum": checksum}
        e2e.update(extra_e2e)
        if relay_info is not None:
            e2e["relay"] = relay_info
            if "value" in relay_info:
                e2e["direct"] = {"value": e2e["value"], "ms_per_step": e2e["ms_per_step"]}
                if relay_on:
                    e2e["value"], e2e["ms_per_step"] = relay_info["value"], relay_info["ms_per_step"]
                    e2e["sample"] += ("; device-to-host copies of the slow-port ranks relayed over NVLink (e2e.relay; e2e.direct = "
                                      "without; e2e.dask / mod03_f32 / ecef measured with the relay on)")
            _capi.set_d2h_relay(None)

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
        else:
            peak, peak_src = FALLBACK_HBM_GBS, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"
        bpp = algorithmic_bytes_per_px(itemsize, cfg)
        achieved = bpp * px_per_rank_step / (kernel_ms_max * 1e-3) / 1e9
        # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture
        # (dram__bytes_read.sum + dram__bytes_write.sum), scaled to this run's granule count
        traffic, traffic_src = None, None
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", "r01_headline_traffic.json")))
            if cap["config_name"] == args.config and cap["dtype"] == args.dtype and args.kernel == "auto":
                traffic = (cap["dram_bytes_read"] + cap["dram_bytes_write"]) * G / cap["granules_per_launch"]
                traffic_src = cap["source"]
        except (OSError, KeyError, ValueError):
            pass
        line = {
            "metric": METRIC if headline else f"interpolated lon/lat pixels/sec ({args.config})",
            "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {
                "workload": (f"CVIIRS 1km->250m, {G} synthetic MODIS granules (2030x1354 {np.dtype(dtype).name} tie points "
                             f"-> 8120x5416 lon+lat each) per GPU per step, one launch per step; BASELINE configs[1] "
                             f"batched x{G}; {G * world} granules whole-job" + (" = configs[3] full day" if G * world == 288 else ""))
                if headline else
                            (f"{args.config}, {G} synthetic MODIS granules ({cfg['tie_rows']}x{cfg['tie_cols']} "
                             f"{np.dtype(dtype).name} tie points -> {cfg['tie_rows'] * cfg['f']}x{cfg['out_cols']} lon+lat each) "
                             f"per GPU per step, one launch per step"),
                "config_name": args.config,
                "granules_per_gpu": G, "kernel": args.kernel,
                "l2": f"working set per step {(bpp * px_per_rank_step) / 1e9:.1f} GB per GPU >> 126 MB L2 (no flush needed)",
                "sharding": "contiguous scan ranges per rank, outputs resident per rank, no collective on the data path",
            },
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes": bpp * px_per_rank_step,
                         "kernel_ms": kernel_ms_max, "algorithmic_bytes_per_px": bpp,
                         "px_per_launch": int(px_per_rank_step)},
            "parity": parity,
        }
        line["config"]["library"] = _capi.library_info()       # what was timed: path + sha256 of the loaded .so
        if sustained is not None:
            sus_value = px_per_rank_step * world / (sustained["ms_per_step"] * 1e-3)
            sustained.update({"value": sus_value, "unit": UNIT,
                              "roofline_frac": bpp * px_per_rank_step / (sustained["ms_per_step"] * 1e-3) / 1e9 / peak})
            line["sustained"] = sustained
            line["roofline"]["sustained"] = sustained            # also here: the driver's summaries keep known keys only
        if world == 1 and headline and not args.no_secondary and args.kernel == "auto":
            # BASELINE configs[2] / [4] and the float32 paths, a few timed launches each (data resident in HBM)
            del d_in[:], outs[:]
            torch.cuda.empty_cache()
            secondary = []
            for name, dn in SECONDARY:
                # 5 km -> 1 km granules are 16 x smaller than 1 km -> 250 m ones: a day of them (the whole batch, 8 times
                # over = 288 granules, half the pixels of the headline launch) - a launch over 36 takes 0.38 ms, of which
                # ~0.05 ms are ramp-up and tail (measured: 67 % of the roofline at 36 granules, 74 % at 144, 76 % at 288)
                five = CONFIGS[name]["coarse"] == 5000
                gs = G if five else min(G, args.secondary_granules)
                base = [np.asarray(a[:gs * TIE_ROWS], np.float64) for a in host_in]
                try:
                    secondary.append(measure_secondary(name, dn, base, gs, device, local_rank, peak, repeat=8 if five else 1))
                except Exception as exc:
                    secondary.append({"config": name, "dtype": dn, "error": repr(exc)})
            line["secondary"] = secondary
            line["roofline"]["secondary"] = secondary
        if e2e is not None:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline and headline:
            try:
                line["cpu_baseline"] = cpu_baseline(dtype)
            except Exception as exc:
                line["cpu_baseline"] = {"error": repr(exc)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    dtype = np.float64 if args.dtype == "f64" else np.float32
    if args.impl == "reference":
        run_reference_arm(args, dtype)
    else:
        run_b200_arm(args, dtype)


if __name__ == "__main__":
    main()
