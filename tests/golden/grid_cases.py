"""Cases of the generic grid interpolators shared by make_grid_golden.py (reference outputs) and the tests."""
import numpy as np

# name -> which reference class, its constructor arguments, tie-grid / fine-grid sizes, where on the globe
CASES = {
    "linear_inside": dict(kind="geogrid", kwargs={}, ny=12, nx=9, my=40, mx=31, margin=0.0, where=(10.0, 45.0)),
    "linear_extrapolate": dict(kind="geogrid", kwargs=dict(bounds_error=False, fill_value=None), ny=12, nx=9, my=40, mx=31,
                               margin=0.08, where=(-75.0, 20.0)),
    "nearest_fill": dict(kind="geogrid", kwargs=dict(method="nearest", bounds_error=False), ny=7, nx=11, my=29, mx=37,
                         margin=0.05, where=(120.0, -50.0)),
    "fill_value_zero": dict(kind="geogrid", kwargs=dict(bounds_error=False, fill_value=0.0), ny=7, nx=11, my=29, mx=37,
                            margin=0.05, where=(120.0, -50.0)),
    "cubic_extrapolate": dict(kind="geogrid", kwargs=dict(method="cubic", bounds_error=False, fill_value=None), ny=10, nx=14,
                              my=33, mx=47, margin=0.04, where=(30.0, 70.0)),
    "slinear_antimeridian": dict(kind="geogrid", kwargs=dict(method="slinear"), ny=8, nx=8, my=25, mx=25, margin=0.0,
                                 where=(179.7, 10.0)),
    "linear_polar": dict(kind="geogrid", kwargs={}, ny=9, nx=9, my=33, mx=33, margin=0.0, where=(40.0, 89.5)),
    "spline_33": dict(kind="geospline", kwargs={}, ny=10, nx=14, my=33, mx=47, margin=0.04, where=(-100.0, 35.0)),
    "spline_13": dict(kind="geospline", kwargs=dict(kx=1, ky=3), ny=10, nx=14, my=33, mx=47, margin=0.0, where=(5.0, -20.0)),
    "spline_11": dict(kind="geospline", kwargs=dict(kx=1, ky=1), ny=6, nx=5, my=21, mx=18, margin=0.1, where=(60.0, 60.0)),
    "plain_arrays_linear": dict(kind="grid", kwargs=dict(bounds_error=False, fill_value=-5.0), ny=6, nx=5, my=21, mx=18,
                                margin=0.1, where=(60.0, 60.0)),
}


def case_inputs(name):
    """(tie_y, tie_x, lon, lat, fine_y, fine_x): a smooth swath-like lon / lat field on a non-uniform grid."""
    spec = CASES[name]
    rng = np.random.default_rng(sorted(CASES).index(name))
    ty = np.cumsum(rng.uniform(3.0, 9.0, spec["ny"]))
    tx = np.cumsum(rng.uniform(2.0, 7.0, spec["nx"]))
    lon0, lat0 = spec["where"]
    yy, xx = np.meshgrid(ty - ty.mean(), tx - tx.mean(), indexing="ij")
    scale = 0.01 / max(np.cos(np.deg2rad(min(abs(lat0), 89.0))), 0.05)
    lon = lon0 + scale * (xx + 0.13 * yy) + 2e-5 * xx * yy
    lat = lat0 + 0.009 * (yy - 0.2 * xx) + 1e-5 * xx * xx
    lon = (lon + 180.0) % 360.0 - 180.0
    lat = np.where(lat > 90.0, 180.0 - lat, lat)
    my, mx, m = spec["my"], spec["mx"], spec["margin"]
    fy = np.linspace(ty[0] - m * (ty[-1] - ty[0]), ty[-1] + m * (ty[-1] - ty[0]), my)
    fx = np.linspace(tx[0] - m * (tx[-1] - tx[0]), tx[-1] + m * (tx[-1] - tx[0]), mx)
    fy[my // 2], fx[mx // 3] = ty[spec["ny"] // 2], tx[spec["nx"] // 3]          # exactly on tie rows / columns too
    return ty, tx, lon, lat, np.sort(fy), np.sort(fx)
