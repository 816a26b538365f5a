"""The oracle's pin against the reference itself, repeated on the GPU box.

``tests/test_oracle_vs_reference.py`` (oracle == oracle/_ref bit for bit) is a CPU test, so the driver's
``-m gpu`` run on the B200 box never re-checked the checker there: the built reference travels with the
snapshot (``oracle/_ref`` is git-ignored but not gpurun-ignored), but nothing loaded it.  This module runs
the very same test functions under the ``gpu`` marker, next to the parity tests that rely on the oracle.
"""
import pytest

import test_oracle_vs_reference as _pin

pytestmark = pytest.mark.gpu

for _name in dir(_pin):
    if _name.startswith("test_"):
        globals()[_name] = getattr(_pin, _name)
del _name
