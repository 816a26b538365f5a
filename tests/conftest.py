"""Shared test plumbing.

``-m "not gpu"`` : oracle vs golden vectors, host logic, C-ABI symbol export (CPU only).
``-m gpu``       : parity tests proper - the CUDA path (through the C ABI) against the oracle.

Only tests may import ``oracle/`` (the CPU restatement of the reference) - see
oracle/gtp_oracle.c.  Nothing here reads /root/reference at run time except
tests/test_oracle_vs_reference.py, which is skipped where oracle/_ref is absent.
"""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (os.path.join(ROOT, "python-geotiepoints_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_available():
    try:
        from geotiepoints_b200 import _capi
        return _capi.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def real_data():
    """The reference's golden file testdata/modis_test_data.h5, re-packed (tests/golden/make_golden.py)."""
    with np.load(os.path.join(GOLDEN, "modis_test_data.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def stress_inputs():
    with np.load(os.path.join(GOLDEN, "stress_inputs.npz")) as z:
        flat = {k: z[k] for k in z.files}
    out = {}
    for key, arr in flat.items():
        kind, name = key.split("/")
        out.setdefault(kind, {})[name] = arr
    return {k: (v["lon"], v["lat"], v["satz"]) for k, v in out.items()}


@pytest.fixture(scope="session")
def reference_digests():
    with open(os.path.join(GOLDEN, "reference_outputs.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden_inputs(real_data, stress_inputs):
    inputs = {"real": (real_data["lon_1km"], real_data["lat_1km"], real_data["satz_1km"])}
    inputs.update(stress_inputs)
    return inputs
