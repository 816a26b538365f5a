"""Build ``libgtp_b200.so`` in-tree:  python -m geotiepoints_b200.build"""
import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")


def build(force=False, verbose=False):
    cmd = ["make", "-C", CSRC] + (["-B"] if force else [])
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode:
        sys.stdout.write(res.stdout)
    if res.returncode:
        raise RuntimeError("building libgtp_b200.so failed")
    return os.path.join(os.path.dirname(CSRC), "libgtp_b200.so")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
