"""The numpy restatement of the LEGACY spline interpolators (oracle/legacy_oracle.py) against the reference.

Three pins:
* the reference's own golden pair ``testdata/test_5_to_1_geoloc_{5km,full}.h5`` to the tolerance of its test
  (``geotiepoints/tests/test_modis.py:41-55``, ``atol=0.05``), re-packed in tests/golden/legacy_golden.npz;
* sampled outputs of the reference's functions (scipy RectBivariateSpline) on synthetic granules, committed in
  the same file by tests/golden/make_legacy_golden.py: 1e-10 deg outside the conditioning bands;
* the reference's functions themselves where oracle/_ref is present (the build container and the GPU box).
"""
import os
import sys

import numpy as np
import pytest

import legacy_oracle
from conftest import GOLDEN, ROOT
from parity import assert_parity

sys.path.insert(0, os.path.join(GOLDEN))
from make_legacy_golden_inputs import GRANULES, legacy_inputs, sample_cols  # noqa: E402

KINDS = ("5km_1km", "1km_500m", "1km_250m")


@pytest.fixture(scope="module")
def golden():
    with np.load(os.path.join(GOLDEN, "legacy_golden.npz")) as z:
        return {k: z[k] for k in z.files}


def test_reference_golden_pair_5km_to_1km(golden):
    """geotiepoints/tests/test_modis.py:41-55"""
    lons, lats = golden["ref5km/lon"] / 1000.0, golden["ref5km/lat"] / 1000.0
    tlons, tlats = legacy_oracle.modis5kmto1km(lons, lats)
    np.testing.assert_allclose(tlons, golden["full1km/lon"] / 1000.0, atol=0.05)
    np.testing.assert_allclose(tlats, golden["full1km/lat"] / 1000.0, atol=0.05)


@pytest.mark.parametrize("granule", GRANULES)
@pytest.mark.parametrize("kind", KINDS)
def test_oracle_matches_committed_reference_samples(golden, kind, granule):
    lon, lat = legacy_inputs(granule, kind)
    got = legacy_oracle.interpolate(lon, lat, kind)
    cols = sample_cols(got[0].shape[1])
    assert_parity(np.ascontiguousarray(got[0][:, cols]), np.ascontiguousarray(got[1][:, cols]),
                  golden[f"{kind}|g{granule}|lon"], golden[f"{kind}|g{granule}|lat"],
                  what=f"legacy oracle {kind} granule {granule} vs reference samples")


@pytest.mark.parametrize("kind", KINDS)
def test_oracle_matches_reference_itself(kind):
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "geotiepoints", "legacy.py")):
        pytest.skip("oracle/_ref/geotiepoints/legacy.py not staged (python oracle/build_ref.py)")
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    from geotiepoints import legacy as ref
    fn = {"5km_1km": ref.modis5kmto1km, "1km_500m": ref.modis1kmto500m, "1km_250m": ref.modis1kmto250m}[kind]
    for granule in (0, 14, 100):
        lon, lat = legacy_inputs(granule, kind, n_scans=3)
        with np.errstate(all="ignore"):
            exp = fn(lon, lat)
        got = legacy_oracle.interpolate(lon, lat, kind)
        assert_parity(*got, *exp, what=f"legacy oracle {kind} granule {granule} vs the reference")


def test_scene_splits():
    """geotiepoints/tests/test_modis.py:27-35"""
    from geotiepoints_b200 import get_scene_splits
    assert list(get_scene_splits(1060, 10, 3)) == [350, 700, 1050]
