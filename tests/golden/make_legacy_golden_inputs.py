"""Inputs of the legacy golden cases (shared by make_legacy_golden.py and the tests; numpy only)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "python-geotiepoints_b200"))
from geotiepoints_b200 import testing  # noqa: E402

GRANULES = (3, 5, 144)          # mid-latitude, polar, antimeridian


def sample_cols(n):
    return np.unique(np.concatenate([np.arange(0, n, 53), np.arange(6), np.arange(n - 6, n)]))


def legacy_inputs(granule, kind, n_scans=2):
    lon, lat, _ = testing.modis_granule(granule, n_scans=n_scans)
    if kind == "5km_1km":
        lon, lat = lon[2::5, 2::5], lat[2::5, 2::5]
    return np.ascontiguousarray(lon), np.ascontiguousarray(lat)
