#!/usr/bin/env python
"""Build the UNMODIFIED reference hot path (and its f64 twin) into ``oracle/_ref/``.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``geotiepoints_b200``; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it.

What it does (sources are read where they lie under /root/reference, never copied
into the repository; every output lands in the git-ignored ``oracle/_ref/``):

* ``oracle/_ref/geotiepoints/``      stock reference: Cython -> C -> gcc -O3 of
  ``_modis_utils.pyx``, ``_modis_interpolator.pyx``, ``_simple_modis_interpolator.pyx``
  (the flags of the reference's ``setup.py:14-17``: ``-O3``, no -march, no fast-math),
  plus the two thin Python wrappers the extensions import at run time
  (``modisinterpolator.py``, ``simple_modis_interpolator.py``) and an EMPTY
  ``__init__.py`` (the reference's own ``__init__`` pulls in the out-of-scope
  legacy spline API).
* ``oracle/_ref/geotiepoints_f64/``  the same tree with the 2-line dtype fix the
  stock CVIIRS path needs to accept float64 at all (SURVEY.md section 0.1:
  ``_modis_interpolator.pyx:148-149`` assign float32 coordinate vectors to
  ``floating`` buffers -> "Buffer dtype mismatch" for float64).  The fix casts the
  two coordinate vectors to the input dtype; the values are small half-integers,
  so the cast is lossless.  Every f64 CVIIRS parity number in this repo that names
  "reference" means this twin, and says so.

* ``oracle/_ref/geotiepoints/{interpolator,geointerpolator}.py`` and ``legacy.py``: the pure-Python
  modules of the LEGACY spline API (``geotiepoints/__init__.py:49-152``,
  ``interpolator.py``, ``geointerpolator.py``), copied as they are; ``legacy.py`` is the reference's
  ``__init__.py`` under another name (the package's ``__init__`` stays empty, see above) with the two
  ``version`` lines at its end left out.  ``oracle/legacy_oracle.py`` is pinned against them.

Run:  python oracle/build_ref.py [--reference /root/reference]
It is a no-op (exit 0) when the reference tree is absent, e.g. on the GPU box,
where the prebuilt ``oracle/_ref`` that travelled with the snapshot is used.
"""
import argparse
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
PYX = ["_modis_utils", "_modis_interpolator", "_simple_modis_interpolator"]
WRAPPERS = ["modisinterpolator.py", "simple_modis_interpolator.py"]
PATCH_FROM = [
    "cdef np.ndarray[floating, ndim=1] x = coords_xy[0]",
    "cdef np.ndarray[floating, ndim=1] y = coords_xy[1]",
]
PATCH_TO = [
    "cdef np.ndarray[floating, ndim=1] x = coords_xy[0].astype(lon1.dtype)",
    "cdef np.ndarray[floating, ndim=1] y = coords_xy[1].astype(lon1.dtype)",
]


def _run(cmd, **kw):
    print("+", " ".join(cmd), flush=True)
    subprocess.check_call(cmd, **kw)


def _stage(ref_pkg, stage_root, pkg_name, patch_f64):
    """Temporary (outside the repo) package tree cython can resolve cimports in."""
    pkg = os.path.join(stage_root, pkg_name)
    os.makedirs(pkg)
    open(os.path.join(pkg, "__init__.py"), "w").close()
    for name in PYX:
        shutil.copy(os.path.join(ref_pkg, name + ".pyx"), pkg)
    shutil.copy(os.path.join(ref_pkg, "_modis_utils.pxd"), pkg)
    for name in WRAPPERS:
        shutil.copy(os.path.join(ref_pkg, name), pkg)
    if patch_f64:
        path = os.path.join(pkg, "_modis_interpolator.pyx")
        src = open(path).read()
        for a, b in zip(PATCH_FROM, PATCH_TO):
            if src.count(a) != 1:
                raise SystemExit(f"f64 patch anchor not found exactly once: {a!r}")
            src = src.replace(a, b)
        open(path, "w").write(src)
    return pkg


def _build_pkg(ref_pkg, pkg_name, patch_f64):
    import numpy as np
    ext = sysconfig.get_config_var("EXT_SUFFIX")
    py_inc = sysconfig.get_paths()["include"]
    dst = os.path.join(OUT, pkg_name)
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    with tempfile.TemporaryDirectory(prefix="gtp_ref_build_") as stage_root:
        pkg = _stage(ref_pkg, stage_root, pkg_name, patch_f64)
        os.makedirs(dst)
        for name in PYX:
            _run([sys.executable, "-m", "cython", "-3", "-X", "freethreading_compatible=True",
                  os.path.join(pkg_name, name + ".pyx")], cwd=stage_root)
            _run(["gcc", "-shared", "-fPIC", "-O3", "-fno-strict-overflow", "-DNDEBUG", "-w",
                  "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
                  "-I", py_inc, "-I", np.get_include(),
                  os.path.join(pkg, name + ".c"), "-o", os.path.join(dst, name + ext), "-lm"])
        for name in WRAPPERS + ["__init__.py"]:
            shutil.copy(os.path.join(pkg, name), dst)
        # keep the generated C of the hot modules beside the .so: it documents the
        # f32 precision model (which expressions are evaluated in double) - SURVEY 8a
        for name in PYX:
            shutil.copy(os.path.join(pkg, name + ".c"), os.path.join(dst, name + ".generated.c"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    args = ap.parse_args()
    ref_pkg = os.path.join(args.reference, "geotiepoints")
    if not os.path.isdir(ref_pkg):
        print(f"reference tree {ref_pkg} absent - keeping any prebuilt oracle/_ref as is")
        return 0
    os.makedirs(OUT, exist_ok=True)
    _build_pkg(ref_pkg, "geotiepoints", patch_f64=False)
    _build_pkg(ref_pkg, "geotiepoints_f64", patch_f64=True)
    # legacy spline API: pure Python, staged beside the stock extensions
    dst = os.path.join(OUT, "geotiepoints")
    for name in ("interpolator.py", "geointerpolator.py"):
        shutil.copy(os.path.join(ref_pkg, name), dst)
    src = open(os.path.join(ref_pkg, "__init__.py")).read()
    tail = "from . import version\n__version__ = version.get_versions()['version']\n"
    if src.count(tail) != 1:
        raise SystemExit("legacy staging: the version lines at the end of geotiepoints/__init__.py moved")
    open(os.path.join(dst, "legacy.py"), "w").write(src.replace(tail, ""))
    with open(os.path.join(OUT, "BUILD_INFO.txt"), "w") as fh:
        import Cython, numpy, scipy
        fh.write(f"reference={args.reference}\ncython={Cython.__version__}\nnumpy={numpy.__version__}\n"
                 f"scipy={scipy.__version__}\ngcc=-O3 (no -march, no -ffast-math)\n"
                 "geotiepoints=stock; geotiepoints_f64=2-line dtype patch at _modis_interpolator.pyx:148-149\n")
    print("built", OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
