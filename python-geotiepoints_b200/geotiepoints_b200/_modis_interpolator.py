"""Operator behind ``modisinterpolator``: satellite-zenith driven ("CVIIRS") interpolation.

Mirrors ``/root/reference/geotiepoints/_modis_interpolator.pyx:17-38`` (``interpolate``)
and ``:93-194`` (``MODISInterpolator``): same arguments, same return value, same errors;
the scan loop (``:251-293``) runs as one CUDA launch through ``gtp_cviirs_interp_*``.
"""
from . import _operator
from ._modis_utils import scanline_mapblocks


@scanline_mapblocks
def interpolate(lon1, lat1, satz1, coarse_resolution=0, fine_resolution=0, coarse_scan_width=0):
    """Interpolate scan-aligned MODIS geolocation arrays (whole scans, full width).

    Inputs are 2-D float32 or float64 arrays (numpy, or torch CUDA tensors for the
    zero-copy device path) of one dtype; returns a new ``(lons, lats)`` tuple.
    """
    if coarse_resolution == 5000 and coarse_scan_width not in (0, 270, 271):
        raise NotImplementedError("Can't interpolate if 5km tiepoints have less than 270 columns.")
    return MODISInterpolator(coarse_resolution, fine_resolution, coarse_scan_width or 0).interpolate(lon1, lat1, satz1)


class MODISInterpolator:
    """Geometry of one (coarse, fine, width) configuration + the launch (``:93-194``)."""

    def __init__(self, coarse_resolution, fine_resolution, coarse_scan_width=0):
        if coarse_resolution == 1000:
            self._coarse_scan_length = 10
            self._coarse_scan_width = 1354
        elif coarse_resolution == 5000:
            self._coarse_scan_length = 2
            self._coarse_scan_width = coarse_scan_width or 271
        else:
            raise ValueError(f"unsupported coarse_resolution {coarse_resolution}")
        if fine_resolution not in (1000, 500, 250) or fine_resolution >= coarse_resolution:
            raise ValueError(f"unsupported fine_resolution {fine_resolution} for coarse_resolution {coarse_resolution}")
        self._coarse_resolution = coarse_resolution
        self._fine_resolution = fine_resolution
        fine_pixels_per_1km = 1000 // fine_resolution
        self._fine_pixels_per_coarse_pixel = fine_pixels_per_1km * (coarse_resolution // 1000)
        self._fine_scan_width = 1354 * fine_pixels_per_1km

    def interpolate_xyz(self, lon1, lat1, satz1):
        """The fine Cartesian coordinates ``(x, y, z)`` of ``_compute_fine_xyz`` (``:346-398``), metres."""
        return self.interpolate(lon1, lat1, satz1, ecef=True)

    def interpolate(self, lon1, lat1, satz1, ecef=False):
        arrays = [lon1, lat1, satz1]
        rows, cols = _operator._check_2d_same_shape(arrays)
        # The reference silently leaves trailing rows uninitialised when the row count is
        # not a whole number of scans and reads out of bounds for a wrong width
        # (SURVEY.md section 7 "hard parts" item 8); this implementation refuses instead.
        if rows % self._coarse_scan_length:
            raise ValueError(f"Input data does not consist of whole scans "
                             f"({self._coarse_scan_length} rows per scan), got {rows} rows.")
        if cols != self._coarse_scan_width:
            raise ValueError(f"Expected {self._coarse_scan_width} tie-point columns, got {cols}.")
        n_scans = rows // self._coarse_scan_length
        out_shape = (rows * self._fine_pixels_per_coarse_pixel, self._fine_scan_width)
        return _operator.run("cviirs", arrays, n_scans, out_shape,
                             (self._coarse_resolution, self._fine_resolution, self._coarse_scan_width), ecef=ecef)
