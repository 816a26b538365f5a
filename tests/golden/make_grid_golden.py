#!/usr/bin/env python
"""Outputs of the reference's GeoGridInterpolator / GeoSplineInterpolator / MultipleGridInterpolator (the staged
copy under oracle/_ref, running scipy) on the cases of grid_cases.py -> tests/golden/grid_golden.npz.

    python oracle/build_ref.py && python tests/golden/make_grid_golden.py

"cubic" is generated with ``solver=scipy.sparse.linalg.spsolve``: the reference's default builds the spline with
an iterative solver (gcrotmk, relative tolerance 1e-5) and is itself only that close to the spline it stands for.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(0, HERE)

from grid_cases import CASES, case_inputs  # noqa: E402


def main():
    import scipy.sparse.linalg as ssl
    from geotiepoints.geointerpolator import GeoGridInterpolator, GeoSplineInterpolator
    from geotiepoints.interpolator import MultipleGridInterpolator
    out = {}
    for name, spec in CASES.items():
        ty, tx, lon, lat, fy, fx = case_inputs(name)
        kind, kwargs = spec["kind"], dict(spec["kwargs"])
        if kwargs.get("method") == "cubic":
            kwargs["solver"] = ssl.spsolve
        if kind == "geogrid":
            res = GeoGridInterpolator((ty, tx), lon, lat, **kwargs).interpolate((fy, fx))
        elif kind == "geospline":
            res = GeoSplineInterpolator((ty, tx), lon, lat, **kwargs).interpolate((fy, fx))
        else:
            res = tuple(MultipleGridInterpolator((ty, tx), lon, lat, **kwargs).interpolate((fy, fx)))
        out[name + "|0"], out[name + "|1"] = np.asarray(res[0]), np.asarray(res[1])
    np.savez_compressed(os.path.join(HERE, "grid_golden.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
