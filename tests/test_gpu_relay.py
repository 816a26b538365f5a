"""The NVLink device-to-host relay of the host entry points (gtp_set_d2h_relay): same bits, another path.
Needs two GPUs with peer access (gpurun --gpus 2); skipped otherwise."""
import numpy as np
import pytest

from geotiepoints_b200 import _capi, ecef, mod03, modisinterpolator, testing

pytestmark = pytest.mark.gpu


def test_relay_gives_identical_results():
    if _capi.device_count() < 2:
        pytest.skip("one GPU only")
    lon, lat, satz = testing.modis_granule(7, n_scans=60)          # several 64 MiB chunks per output
    direct = modisinterpolator.modis_1km_to_250m(lon, lat, satz)
    direct = (np.array(direct[0]), np.array(direct[1]))
    try:
        _capi.set_d2h_relay(1)
        assert _capi.lib().gtp_get_d2h_relay() == 1
        for _ in range(2):                                          # second call: cached pipe, staging buffers reused
            relayed = modisinterpolator.modis_1km_to_250m(lon, lat, satz)
            assert np.array_equal(relayed[0], direct[0]) and np.array_equal(relayed[1], direct[1])
        xyz_r = ecef.modis_1km_to_250m(lon, lat, satz)
        zen = np.round(satz / 0.01).astype(np.int16)
        m_r = mod03.modis_1km_to_250m(lon.astype(np.float32), lat.astype(np.float32), zen)
        xyz_r, m_r = [np.array(a) for a in xyz_r], [np.array(a) for a in m_r]
    finally:
        _capi.set_d2h_relay(None)
    assert _capi.lib().gtp_get_d2h_relay() == -1
    again = modisinterpolator.modis_1km_to_250m(lon, lat, satz)
    assert np.array_equal(again[0], direct[0]) and np.array_equal(again[1], direct[1])
    xyz_d = ecef.modis_1km_to_250m(lon, lat, satz)
    m_d = mod03.modis_1km_to_250m(lon.astype(np.float32), lat.astype(np.float32), zen)
    assert all(np.array_equal(a, b) for a, b in zip(xyz_r, xyz_d)) and all(np.array_equal(a, b) for a, b in zip(m_r, m_d))
    with pytest.raises(ValueError):
        _capi.set_d2h_relay(1000)
