#!/usr/bin/env python
"""Experimental builds of libgtp_b200.so, kept OUT of the product sources.

The power / A-B experiments of DESIGN.md 4.5 / 4.6 need kernels that leave out the per-pixel math, the HBM
stores or the register parking.  They used to be ``#ifdef`` switches inside csrc/; a "math-less" build must
not be one ``-D`` away from the shipped kernel, so they are textual patches applied to a COPY of csrc/ under
build/ab/<name>/ and built into build/ab/libgtp_b200_<name>.so.  Load one with GTP_B200_LIB=... (the loader
warns, bench.py prints the path and sha256 it timed).

    python tools/exp_variants.py no_math | no_store | no_park
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "python-geotiepoints_b200", "geotiepoints_b200", "csrc")

PATCHES = {
    # keep the HBM writes, drop the per-pixel math
    "no_math": [("gtp_fast.cu",
                 "            if (MODE == kWide) pixel_eval_wide(q, rc, a_col[k], vlon[k], vlat[k]);\n"
                 "            else pixel_eval<MODE, WRAP>(q, rc, a_col[k], vlon[k], vlat[k]);\n",
                 "            vlon[k] = a_col[k] + rc.Ve; vlat[k] = a_col[k] - rc.Ue;\n")],
    # keep the math, drop the HBM writes
    "no_store": [("gtp_fast_io.cuh",
                  "template <>\n__device__ __forceinline__ void store_px<4>(double* p, const double (&v)[4])\n{\n",
                  "template <>\n__device__ __forceinline__ void store_px<4>(double* p, const double (&v)[4])\n{\n"
                  "    if (v[0] != 1234.56789) return;\n")],
    # no per-thread shared-memory parking of the set-up values in the 250 m kernel
    "no_park": [("gtp_fast.cu", "    constexpr bool kPark = (P == 4);\n", "    constexpr bool kPark = false;\n")],
}


def build(name):
    out = os.path.join(ROOT, "build", "ab", name)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    for f in os.listdir(CSRC):
        if f.endswith((".cu", ".cuh", ".h")):
            shutil.copy(os.path.join(CSRC, f), out)
    for fname, old, new in PATCHES[name]:
        path = os.path.join(out, fname)
        text = open(path).read()
        if old not in text:
            raise SystemExit(f"{name}: the text to patch is no longer in {fname}")
        open(path, "w").write(text.replace(old, new, 1))
    for f in os.listdir(out):
        path = os.path.join(out, f)
        text = open(path).read().replace("../../../include/gtp_b200.h", os.path.join(ROOT, "include", "gtp_b200.h"))
        open(path, "w").write(text)
    flags = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"]
    objs = []
    for f in sorted(os.listdir(out)):
        if f.endswith(".cu"):
            extra = ["-fmad=false"] if f in ("gtp_generic.cu", "gtp_fast_f32.cu") else []
            obj = os.path.join(out, f[:-3] + ".o")
            subprocess.check_call(["nvcc", *flags, *extra, "-c", os.path.join(out, f), "-o", obj])
            objs.append(obj)
    lib = os.path.join(ROOT, "build", "ab", f"libgtp_b200_{name}.so")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib, *objs])
    return lib


if __name__ == "__main__":
    if len(sys.argv) != 2 or sys.argv[1] not in PATCHES:
        raise SystemExit(__doc__)
    print(build(sys.argv[1]))
