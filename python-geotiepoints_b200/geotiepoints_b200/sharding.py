"""Scan-range sharding of a batch of granules over the GPUs of one box.

Scans are independent (no scan reads another scan's tie points:
``/root/reference/geotiepoints/_modis_interpolator.pyx:272-293``), so a batch flattened to
``n_scans`` scans is cut into contiguous ranges, one per rank, with no exchange step:
there is no collective on the data path.  Outputs stay resident on the owning GPU.
"""


def scan_range(rank, world_size, n_scans):
    """Half-open scan range ``[lo, hi)`` owned by ``rank``."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return (rank * n_scans) // world_size, ((rank + 1) * n_scans) // world_size


def granule_range(rank, world_size, n_granules):
    """Half-open range of whole granules owned by ``rank``."""
    return scan_range(rank, world_size, n_granules)


def max_over_ranks(value, dist=None):
    """Max of a per-rank scalar (elapsed device time) - the only collective the bench uses."""
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
