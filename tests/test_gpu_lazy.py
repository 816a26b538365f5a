"""SURVEY.md 8(f) row 1 on the GPU: lazy (dask-shaped) inputs through ``scanline_mapblocks`` into the CUDA
operators, compared with the oracle.

The reference's real entry point is dask / xarray (``geotiepoints/tests/test_modisinterpolator.py:22-27``
``chunks=4096``; laziness asserted ``:127-129``; ``test_simple_modis_interpolator.py:47-53`` for the partial
scan).  dask and xarray are not installed in this image, so the block mapping is driven by
``tests/fake_dask.py``, which models exactly what the adapter touches and counts computes like the
reference's ``CustomScheduler``; the same tests run against real dask where it can be imported.
"""
import types

import numpy as np
import pytest

import fake_dask
import gtp_oracle
from geotiepoints_b200 import _modis_utils, _operator, ecef, mod03, modisinterpolator, simple_modis_interpolator, testing
from parity import assert_parity, assert_xyz_parity

pytestmark = pytest.mark.gpu


class FakeDataArray:
    """The two attributes of xarray.DataArray the adapter uses (_modis_utils.pyx:132-133, :155-162)."""
    def __init__(self, data, dims):
        self.data, self.dims = data, tuple(dims)
        self.ndim, self.shape, self.dtype = data.ndim, data.shape, data.dtype

    def compute(self):
        return self.data.compute()


@pytest.fixture
def lazy(monkeypatch):
    mod = types.SimpleNamespace(map_blocks=fake_dask.map_blocks, Array=fake_dask.Array)
    monkeypatch.setattr(_modis_utils, "da", mod)
    monkeypatch.setattr(_modis_utils, "xr", types.SimpleNamespace(DataArray=FakeDataArray))
    fake_dask.COMPUTES["n"] = 0
    fake_dask.BLOCK_SHAPES.clear()
    return mod


def _chunked(arrays, chunks):
    return [fake_dask.from_array(a, chunks) for a in arrays]


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["float64", "float32"])
@pytest.mark.parametrize("chunks", [(25, 700), (4096, 4096), (40, 1354)], ids=["misaligned", "one_block", "aligned"])
def test_lazy_cviirs_1km_to_250m(lazy, dtype, chunks):
    lon, lat, satz = testing.modis_granule(3, n_scans=7, dtype=dtype)
    d = _chunked((lon, lat, satz), chunks)
    lons, lats = modisinterpolator.modis_1km_to_250m(*d)
    assert fake_dask.COMPUTES["n"] == 0                      # tests/test_modisinterpolator.py:127-129
    assert lons.shape == (280, 5416) and lons.dtype == dtype
    got = (lons.compute(), lats.compute())
    assert all(s[0] % 10 == 0 and s[1] == 1354 for s in fake_dask.BLOCK_SHAPES)      # whole scans x full width
    exp = gtp_oracle.cviirs(lon, lat, satz, 1000, 250)
    assert_parity(*got, *exp, what=f"lazy cviirs 1km->250m {np.dtype(dtype).name} chunks {chunks}")
    eager = modisinterpolator.modis_1km_to_250m(lon, lat, satz)
    assert np.array_equal(got[0], eager[0]) and np.array_equal(got[1], eager[1])     # block mapping changes nothing


def test_lazy_5km_and_simple(lazy):
    lon, lat, satz = testing.modis_granule(14, n_scans=20, dtype=np.float32)
    lon5, lat5, satz5 = testing.subsample_5km(lon, lat, satz, 271)
    d5 = _chunked((lon5, lat5, satz5), (14, 100))
    lons, lats = modisinterpolator.modis_5km_to_1km(*d5)
    assert fake_dask.COMPUTES["n"] == 0 and lons.shape == (200, 1354)
    with np.errstate(all="ignore"):
        exp5 = gtp_oracle.cviirs(lon5, lat5, satz5, 5000, 1000, 271)
    assert_parity(lons.compute(), lats.compute(), *exp5, what="lazy cviirs 5km->1km float32")
    for fine, fn in ((250, simple_modis_interpolator.modis_1km_to_250m), (500, simple_modis_interpolator.modis_1km_to_500m)):
        d = _chunked((lon, lat), (33, 1354))
        lons, lats = fn(*d)
        exp = gtp_oracle.simple(lon, lat, 1000, fine)
        assert_parity(lons.compute(), lats.compute(), *exp, what=f"lazy simple 1km->{fine} float32")


def test_lazy_partial_scan_is_refused(lazy):
    """tests/test_simple_modis_interpolator.py:47-53: dropping the last row must raise, before any compute."""
    lon, lat, satz = testing.modis_granule(3, n_scans=2, dtype=np.float32)
    d = _chunked((lon[:-1], lat[:-1], satz[:-1]), 4096)
    with pytest.raises(ValueError, match=r"does not consist of whole scans \(10 rows per scan\)"):
        simple_modis_interpolator.modis_1km_to_250m(d[0], d[1])
    with pytest.raises(ValueError, match=r"does not consist of whole scans \(10 rows per scan\)"):
        modisinterpolator.modis_1km_to_250m(*d)
    assert fake_dask.COMPUTES["n"] == 0


def test_lazy_dataarrays_are_rewrapped(lazy):
    lon, lat, satz = testing.modis_granule(0, n_scans=3, dtype=np.float32)
    xa = [FakeDataArray(a, ("y", "x")) for a in _chunked((lon, lat, satz), 4096)]
    res = modisinterpolator.modis_1km_to_500m(*xa)
    assert fake_dask.COMPUTES["n"] == 0
    assert len(res) == 2 and all(isinstance(r, FakeDataArray) and r.dims == ("y", "x") for r in res)
    exp = gtp_oracle.cviirs(lon, lat, satz, 1000, 500)
    assert_parity(res[0].compute(), res[1].compute(), *exp, what="lazy DataArray cviirs 1km->500m float32")


def test_lazy_blocks_on_the_gpu_stay_there(lazy):
    """torch CUDA tensors as blocks: the operator's outputs are planes of one device buffer, the adapter's
    stacking is a view of it, nothing crosses PCIe."""
    import torch
    lon, lat, satz = testing.modis_granule(5, n_scans=6)
    dev = [torch.from_numpy(a).cuda() for a in (lon, lat, satz)]
    d = _chunked(dev, (30, 1354))
    lons, lats = modisinterpolator.modis_1km_to_250m(*d)
    assert fake_dask.COMPUTES["n"] == 0
    got_lon, got_lat = lons.compute(), lats.compute()
    assert got_lon.is_cuda and got_lat.is_cuda and got_lon.dtype == torch.float64
    with np.errstate(all="ignore"):
        exp = gtp_oracle.cviirs(lon, lat, satz, 1000, 250)
    assert_parity(got_lon.cpu().numpy(), got_lat.cpu().numpy(), *exp, what="lazy torch blocks cviirs 1km->250m")
    out = modisinterpolator.modis_1km_to_250m(*dev)
    stacked = _operator.stack(out)
    assert stacked.data_ptr() == out[0].data_ptr() and stacked.shape == (2, 240, 5416)
    assert stacked.untyped_storage().data_ptr() == out[1].untyped_storage().data_ptr()


def test_stacking_numpy_outputs_is_a_view():
    lon, lat, satz = testing.modis_granule(0, n_scans=2)
    out = modisinterpolator.modis_1km_to_250m(lon, lat, satz)
    stacked = _operator.stack(out)
    assert stacked.shape == (2, 80, 5416)
    assert np.shares_memory(stacked, out[0]) and np.shares_memory(stacked, out[1])
    assert stacked.ctypes.data == out[0].ctypes.data
    assert np.array_equal(stacked[0], out[0]) and np.array_equal(stacked[1], out[1])
    separate = (out[0].copy(), out[1].copy())                # arrays that are NOT planes of one block: copied
    assert not np.shares_memory(_operator.stack(separate), separate[0])


def test_lazy_ecef_and_mod03(lazy):
    lon, lat, satz = testing.modis_granule(4, n_scans=5)
    d = _chunked((lon, lat, satz), (20, 1354))
    xyz = ecef.modis_1km_to_250m(*d)
    assert fake_dask.COMPUTES["n"] == 0 and len(xyz) == 3 and xyz[0].shape == (200, 5416)
    assert_xyz_parity(tuple(a.compute() for a in xyz), gtp_oracle.cviirs_xyz(lon, lat, satz, 1000, 250), what="lazy ecef")
    lon32, lat32 = lon.astype(np.float32), lat.astype(np.float32)
    zen = np.round(satz / 0.01).astype(np.int16)
    d = _chunked((lon32, lat32, zen), (17, 1354))
    fake_dask.COMPUTES["n"] = 0
    lons, lats = mod03.modis_1km_to_250m(*d)                 # int16 zenith: scale 0.01 by default
    assert fake_dask.COMPUTES["n"] == 0 and lons.dtype == np.float32
    exp = gtp_oracle.cviirs(lon32.astype(np.float64), lat32.astype(np.float64), zen.astype(np.float64) * 0.01, 1000, 250)
    got = (lons.compute(), lats.compute())
    assert got[0].dtype == np.float32
    assert_parity(*got, *(a.astype(np.float32) for a in exp), what="lazy mod03 1km->250m")


def test_real_dask_and_xarray():
    """The same with the real libraries, where they are installed (they are not in this image)."""
    da = pytest.importorskip("dask.array")
    xr = pytest.importorskip("xarray")
    import dask
    lon, lat, satz = testing.modis_granule(3, n_scans=8, dtype=np.float32)
    xa = [xr.DataArray(da.from_array(a, chunks=(25, 700)), dims=("y", "x")) for a in (lon, lat, satz)]
    lons, lats = modisinterpolator.modis_1km_to_250m(*xa)
    assert isinstance(lons, xr.DataArray) and lons.dims == ("y", "x") and lons.shape == (320, 5416)
    with dask.config.set(scheduler="threads"):
        got = dask.compute(lons, lats)
    exp = gtp_oracle.cviirs(lon, lat, satz, 1000, 250)
    assert_parity(np.asarray(got[0]), np.asarray(got[1]), *exp, what="real dask + xarray cviirs 1km->250m float32")
