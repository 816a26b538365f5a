"""Case tables shared by the CPU and GPU test modules."""
import numpy as np

from geotiepoints_b200 import testing

CVIIRS_CONFIGS = [  # (coarse, fine, width)
    (1000, 250, 0), (1000, 500, 0),
    (5000, 1000, 271), (5000, 500, 271), (5000, 250, 271),
    (5000, 1000, 270), (5000, 500, 270), (5000, 250, 270),
]
SIMPLE_CONFIGS = [(1000, 250), (1000, 500)]
INPUT_NAMES = ["real", "antimeridian", "pole", "south_pole", "prime_meridian", "lat_switch",
               "symmetric_satz", "zeros", "nan_holes"]
DTYPES = [np.float32, np.float64]


def case_id(inp, algo, coarse, fine, width, dtype):
    return f"{inp}|{algo}|{coarse}->{fine}|w{width}|{np.dtype(dtype).name}"


def cviirs_inputs(arrs32, coarse, width, dtype):
    lon, lat, satz = (a.astype(dtype) for a in arrs32)
    if coarse == 5000:
        return testing.subsample_5km(lon, lat, satz, width)
    return lon, lat, satz
