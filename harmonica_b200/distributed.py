"""
One-process-per-GPU sharding of the pairwise forward models over
``torch.distributed`` (NCCL on GPUs; gloo in the CPU test-suite).

The path shards without any data-path exchange when observers are split
(disjoint output slices; the only communication is the final gather of the
result) and needs ONE collective when sources are split: a float64 sum of the
per-rank partial fields. The in-process multi-GPU path of the C
library (``hb200_init`` with several devices) does the same with peer copies.
"""

import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous, balanced [lo, hi) of ``n`` units for ``rank`` of ``world``."""
    return n * rank // world, n * (rank + 1) // world


def _device_for(backend):
    import torch  # noqa: PLC0415

    return torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else "cpu"


def observer_sharded(compute, coordinates, n_fields=1, group=None):
    """
    ``compute(sub_coordinates) -> array (n_fields, n_local)`` on this rank's
    observer slice; returns the full ``(n_fields, n_obs)`` result on every rank.
    """
    import torch  # noqa: PLC0415
    import torch.distributed as dist  # noqa: PLC0415

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    coords = tuple(np.ascontiguousarray(np.asarray(c, dtype=np.float64).ravel()) for c in coordinates[:3])
    n_obs = coords[0].size
    lo, hi = shard_bounds(n_obs, rank, world)
    local = np.asarray(compute(tuple(c[lo:hi] for c in coords)), dtype=np.float64).reshape(n_fields, hi - lo)
    device = _device_for(dist.get_backend(group))
    # equal-sized slots so that all_gather works for ragged shards
    slot = (n_obs + world - 1) // world
    send = torch.zeros((n_fields, slot), dtype=torch.float64, device=device)
    send[:, : hi - lo] = torch.from_numpy(local).to(device)
    recv = [torch.empty_like(send) for _ in range(world)]
    dist.all_gather(recv, send, group=group)
    out = np.empty((n_fields, n_obs), dtype=np.float64)
    for r in range(world):
        a, b = shard_bounds(n_obs, r, world)
        out[:, a:b] = recv[r][:, : b - a].cpu().numpy()
    return out


def source_sharded(compute_partial, n_sources, n_obs, n_fields=1, group=None):
    """
    ``compute_partial(lo, hi) -> array (n_fields, n_obs)``: the field of sources
    ``[lo, hi)`` on ALL observers (linear units). Returns the all-reduced sum.
    """
    import torch  # noqa: PLC0415
    import torch.distributed as dist  # noqa: PLC0415

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(n_sources, rank, world)
    part = np.asarray(compute_partial(lo, hi), dtype=np.float64).reshape(n_fields, n_obs)
    t = torch.from_numpy(np.ascontiguousarray(part)).to(_device_for(dist.get_backend(group)))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy()
