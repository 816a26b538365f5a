"""CPU restatement of the reference's generic rectilinear-grid interpolators (numpy only).

TEST INFRASTRUCTURE: only ``tests/``, ``__graft_entry__.smoke()`` and the CPU legs of ``bench.py`` may
import it; the product package never does.

Restates what ``SingleGridInterpolator`` / ``SingleSplineInterpolator``
(``/root/reference/geotiepoints/interpolator.py:232-340``) and ``GeoGridInterpolator`` /
``GeoSplineInterpolator`` (``geointerpolator.py:76-113``) compute.  The arithmetic itself lives in a third-party
dependency that is not under ``/root/reference``: ``scipy.interpolate.RegularGridInterpolator`` and
``RectBivariateSpline`` (unpinned: ``requirements.txt`` ``scipy>=0.14``; installed here: 1.18.1).  Their
published algorithms, per axis of a tensor product:

* ``"nearest"``: interval ``i`` with ``grid[i] <= x < grid[i + 1]`` (clipped to the grid), normalised distance
  ``u = (x - grid[i]) / (grid[i + 1] - grid[i])``, value of node ``i`` if ``u <= 0.5`` else of node ``i + 1``;
* ``"linear"`` (also ``"slinear"`` and spline order 1): ``v_i (1 - u) + v_{i+1} u`` with the same ``i``, ``u``;
  ``u`` outside [0, 1] is the linear extrapolation scipy performs for ``fill_value=None``;
* ``"cubic"`` (spline order 3): the interpolating cubic spline with not-a-knot end conditions
  (``make_interp_spline(k=3)``'s default; FITPACK's knot choice for ``s = 0``), its end pieces continued outside
  the grid;
* out of the grid: extrapolation (``fill_value=None``), a fill value (``bounds_error=False``), or - FITPACK - the
  arguments clamped to the grid.

Pinned against the reference classes themselves (``oracle/_ref/geotiepoints``, running scipy) and against the
expected vectors of ``geotiepoints/tests/test_geointerpolator.py:187-278`` in ``tests/test_grid_oracle.py``.
"""
import numpy as np

try:
    from .legacy_oracle import lonlat2xyz, notaknot_second_derivatives, xyz2lonlat
except ImportError:                     # tests put oracle/ itself on sys.path
    from legacy_oracle import lonlat2xyz, notaknot_second_derivatives, xyz2lonlat

NEAREST, LINEAR, CUBIC = 0, 1, 3
EXTRAPOLATE, FILL, CLAMP = 0, 1, 2


def locate(tie, fine, clamp=False):
    """Interval index, offset into it, its length, and the out-of-range mask of every fine coordinate."""
    tie, fine = np.asarray(tie, np.float64), np.asarray(fine, np.float64)
    outside = (fine < tie[0]) | (fine > tie[-1])
    x = np.clip(fine, tie[0], tie[-1]) if clamp else fine
    idx = np.clip(np.searchsorted(tie, x, side="right") - 1, 0, tie.size - 2)
    return idx, x - tie[idx], tie[idx + 1] - tie[idx], outside


def interp_axis(tie, values, fine, method, clamp=False):
    """1-D scheme along the LAST axis of `values` (shape (..., tie.size)) at `fine`."""
    tie = np.asarray(tie, np.float64)
    idx, d, h, _ = locate(tie, fine, clamp)
    v0, v1 = values[..., idx], values[..., idx + 1]
    if method == NEAREST:
        return np.where(d / h <= 0.5, v0, v1)
    if method == LINEAR:
        u = d / h
        return v0 * (1.0 - u) + v1 * u
    if method == CUBIC:
        flat = values.reshape(-1, tie.size)
        M = notaknot_second_derivatives(tie, flat).reshape(values.shape)
        m0, m1 = M[..., idx], M[..., idx + 1]
        b = (v1 - v0) / h - h * (2.0 * m0 + m1) / 6.0
        return v0 + d * (b + d * (0.5 * m0 + d * ((m1 - m0) / (6.0 * h))))
    raise ValueError(method)


def grid_interpolate(tie_y, tie_x, values, fine_y, fine_x, method_y=LINEAR, method_x=LINEAR, oob=EXTRAPOLATE,
                     fill=np.nan):
    """values (ny, nx) on tie_y x tie_x  ->  (len(fine_y), len(fine_x)): along x first, then along y."""
    values = np.asarray(values, np.float64)
    clamp = oob == CLAMP
    rows = interp_axis(tie_x, values, fine_x, method_x, clamp)                 # (ny, mx)
    out = interp_axis(tie_y, rows.T, fine_y, method_y, clamp).T                # (my, mx)
    if oob == FILL:
        oy, ox = locate(tie_y, fine_y)[3], locate(tie_x, fine_x)[3]
        out = np.where(oy[:, None] | ox[None, :], fill, out)
    return out


def geogrid_interpolate(tie_y, tie_x, lon, lat, fine_y, fine_x, **kwargs):
    """GeoGridInterpolator / GeoSplineInterpolator: the scheme on Cartesian coordinates, then xyz2lonlat."""
    with np.errstate(all="ignore"):
        xyz = lonlat2xyz(np.asarray(lon, np.float64), np.asarray(lat, np.float64))
        fine = [grid_interpolate(tie_y, tie_x, comp, fine_y, fine_x, **kwargs) for comp in xyz]
        return xyz2lonlat(*fine)
