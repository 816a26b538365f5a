"""World-size-2 test of the multi-rank plumbing over gloo (CPU): scan sharding with no
data-path collective, and the max-over-ranks reduction the bench reports timing with."""
import os
import subprocess
import sys
import textwrap

from conftest import ROOT

WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, os.path.join({root!r}, "python-geotiepoints_b200"))
    sys.path.insert(0, os.path.join({root!r}, "oracle"))
    import numpy as np
    import torch
    import torch.distributed as dist
    from geotiepoints_b200 import sharding, testing
    import gtp_oracle

    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n_scans = 5
    lon, lat, satz = testing.modis_granule(2, n_scans=n_scans, dtype=np.float32)
    lo, hi = sharding.scan_range(rank, world, n_scans)
    # each rank interpolates only its scans (stand-in operator: the CPU oracle)
    mine = gtp_oracle.cviirs(lon[lo*10:hi*10], lat[lo*10:hi*10], satz[lo*10:hi*10], 1000, 500)
    whole = gtp_oracle.cviirs(lon, lat, satz, 1000, 500)
    ok = all(np.array_equal(m, w[lo*20:hi*20]) for m, w in zip(mine, whole))
    ranges = [None] * world
    dist.all_gather_object(ranges, (lo, hi))
    slowest = sharding.max_over_ranks(1.0 + rank, dist)
    dist.barrier()
    if rank == 0:
        print(json.dumps({{"ok": ok, "ranges": ranges, "slowest": slowest}}))
    else:
        assert ok
    dist.destroy_process_group()
""")


def test_two_ranks_shard_scans_without_exchange(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    res = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
        capture_output=True, text=True, env=env, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    import json
    line = [l for l in res.stdout.splitlines() if l.startswith("{")][-1]
    out = json.loads(line)
    assert out["ok"] is True
    assert out["ranges"] == [[0, 2], [2, 5]]
    assert out["slowest"] == 2.0
