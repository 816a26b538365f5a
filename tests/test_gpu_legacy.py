"""SURVEY.md 8(f) row 4 on the GPU: the legacy spline interpolators (csrc/gtp_legacy.cu) through the C ABI against
the numpy restatement (oracle/legacy_oracle.py, pinned against the reference in tests/test_legacy_oracle.py), the
reference's own golden pair, and the reference's functions themselves where oracle/_ref travelled along."""
import os
import sys

import numpy as np
import pytest

import legacy_oracle
from conftest import GOLDEN, ROOT
from geotiepoints_b200 import legacy, testing
from parity import assert_parity

sys.path.insert(0, GOLDEN)
from make_legacy_golden_inputs import legacy_inputs  # noqa: E402

pytestmark = pytest.mark.gpu

FUNCS = {"5km_1km": legacy.modis5kmto1km, "1km_500m": legacy.modis1kmto500m, "1km_250m": legacy.modis1kmto250m}


@pytest.mark.parametrize("granule", [0, 3, 5, 14, 144, 287])
@pytest.mark.parametrize("kind", list(FUNCS))
def test_legacy_matches_oracle(kind, granule):
    lon, lat = legacy_inputs(granule, kind, n_scans=5)
    got = FUNCS[kind](lon, lat)
    exp = legacy_oracle.interpolate(lon, lat, kind)
    assert got[0].dtype == np.float64 and got[0].shape == exp[0].shape
    assert_parity(*got, *exp, what=f"legacy {kind} granule {granule}")


def test_reference_golden_pair_5km_to_1km():
    """geotiepoints/tests/test_modis.py:41-55 (int32 millidegrees in, atol = 0.05)"""
    with np.load(os.path.join(GOLDEN, "legacy_golden.npz")) as z:
        lons, lats = z["ref5km/lon"] / 1000.0, z["ref5km/lat"] / 1000.0
        glons, glats = z["full1km/lon"] / 1000.0, z["full1km/lat"] / 1000.0
    tlons, tlats = legacy.modis5kmto1km(lons, lats)
    np.testing.assert_allclose(tlons, glons, atol=0.05)
    np.testing.assert_allclose(tlats, glats, atol=0.05)


def test_zero_section_1km_to_250m():
    """geotiepoints/tests/test_modis.py:57-75: testdata/250m_lonlat_section_{input,result}.h5 hold zeros only
    (SURVEY.md 0.2): lon = lat = 0 must come back as exactly 0, with and without `cores`."""
    zeros = np.zeros((30, 1354), np.int32)
    for kwargs in ({}, {"cores": 4}):
        tlons, tlats = legacy.modis1kmto250m(zeros / 1000.0, zeros / 1000.0, **kwargs)
        assert tlons.shape == (120, 5416)
        np.testing.assert_allclose(tlons, 0.0, atol=0.05)
        np.testing.assert_allclose(tlats, 0.0, atol=0.05)


def test_device_path_and_input_dtypes():
    import torch
    lon, lat = legacy_inputs(3, "1km_250m", n_scans=3)
    host = legacy.modis1kmto250m(lon, lat)
    dev = legacy.modis1kmto250m(torch.from_numpy(lon).cuda(), torch.from_numpy(lat).cuda())
    assert dev[0].is_cuda and dev[0].dtype == torch.float64
    assert np.array_equal(dev[0].cpu().numpy(), host[0]) and np.array_equal(dev[1].cpu().numpy(), host[1])
    f32 = legacy.modis1kmto250m(lon.astype(np.float32), lat.astype(np.float32))
    assert f32[0].dtype == np.float64 and np.max(np.abs(f32[0] - host[0])) < 1e-4      # widened, not float32 arithmetic
    with pytest.raises(ValueError, match="whole scans"):
        legacy.modis1kmto250m(lon[:-1], lat[:-1])


@pytest.mark.parametrize("kind", list(FUNCS))
def test_legacy_matches_reference_itself(kind):
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_dir, "geotiepoints", "legacy.py")):
        pytest.skip("oracle/_ref/geotiepoints/legacy.py not staged")
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    from geotiepoints import legacy as ref
    fn = {"5km_1km": ref.modis5kmto1km, "1km_500m": ref.modis1kmto500m, "1km_250m": ref.modis1kmto250m}[kind]
    lon, lat = legacy_inputs(41, kind, n_scans=4)
    with np.errstate(all="ignore"):
        exp = fn(lon, lat)
    assert_parity(*FUNCS[kind](lon, lat), *exp, what=f"legacy {kind} vs the reference (scipy RectBivariateSpline)")


def test_full_granule_batch_is_scan_independent():
    """Scans are independent chunks (chunk_size = one scan): a granule computed at once equals its halves."""
    lon, lat, _ = testing.modis_granule(9, n_scans=40)
    whole = legacy.modis1kmto500m(lon, lat)
    a, b = legacy.modis1kmto500m(lon[:170], lat[:170]), legacy.modis1kmto500m(lon[170:], lat[170:])
    assert np.array_equal(whole[0], np.vstack((a[0], b[0]))) and np.array_equal(whole[1], np.vstack((a[1], b[1])))
