"""
world_size-2 `gloo` test of the N>1 sharding logic (harmonica_b200/distributed.py)
on CPU. The per-rank compute is the oracle here (there is no GPU); on a GPU box
the same helpers wrap harmonica_b200 calls with one process per GPU over NCCL.
"""

import os
import socket

import numpy as np
import pytest

from _common import TOL, config1, max_rel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, queue):
    import sys

    import torch.distributed as dist

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [os.path.dirname(here), os.path.join(os.path.dirname(here), "oracle"), here]
    import oracle as O
    from harmonica_b200 import distributed as hbd

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    coords, prisms, density = config1(301, 203, seed=9)  # ragged: 203 and 301 are odd
    full = hbd.observer_sharded(
        lambda sub: np.stack([O.prism_gravity(sub, prisms, density, f) for f in ("g_z", "g_zz")]),
        coords, n_fields=2)
    pts = (prisms[:, 0], prisms[:, 2], prisms[:, 4])
    part = hbd.source_sharded(
        lambda lo, hi: O.eqs_predict(coords, tuple(p[lo:hi] for p in pts), density[lo:hi]),
        n_sources=301, n_obs=203)
    lo, hi = hbd.shard_bounds(203, rank, world)
    # the collectives ShardedJob uses on the device tensors: gather to dst only / reduce to dst only
    import torch

    local = torch.from_numpy(np.arange(3 * (hi - lo), dtype=np.float64).reshape(3, hi - lo) + 1000 * rank)
    to0 = hbd.gather_observer_slices(local, 203, dst=0)
    to1 = hbd.reduce_source_partials(torch.full((2, 5), float(rank + 1), dtype=torch.float64), dst=1)
    even = hbd.gather_observer_slices(torch.full((1, 4), float(rank), dtype=torch.float64), 8, dst=None)
    extra = (None if to0 is None else to0.numpy(), None if to1 is None else to1.numpy(), even.numpy())
    queue.put((rank, full, part, (lo, hi), extra))
    dist.barrier()
    dist.destroy_process_group()


def test_observer_and_source_sharding_world2():
    import torch.multiprocessing as mp

    import oracle as O

    ctx = mp.get_context("spawn")
    queue, port = ctx.Queue(), _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, queue)) for r in range(2)]
    for p in procs:
        p.start()
    results = [queue.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    coords, prisms, density = config1(301, 203, seed=9)
    want = np.stack([O.prism_gravity(coords, prisms, density, f) for f in ("g_z", "g_zz")])
    pts = (prisms[:, 0], prisms[:, 2], prisms[:, 4])
    want_eqs = O.eqs_predict(coords, pts, density)
    bounds = sorted(r[3] for r in results)
    assert bounds == [(0, 101), (101, 203)]
    for rank, full, part, _, extra in results:
        np.testing.assert_array_equal(full, want)          # disjoint slices: bit-exact
        assert max_rel(part[0], want_eqs) <= TOL            # re-associated sum
        to0, to1, even = extra
        if rank == 0:  # ragged gather on dst: rank 0 owns 101 columns, rank 1 owns 102
            assert to1 is None and to0.shape == (3, 203)
            np.testing.assert_array_equal(to0[:, :101], np.arange(303.0).reshape(3, 101))
            np.testing.assert_array_equal(to0[:, 101:], np.arange(306.0).reshape(3, 102) + 1000)
        else:
            assert to0 is None
            np.testing.assert_array_equal(to1, np.full((2, 5), 3.0))
        np.testing.assert_array_equal(even, [[0, 0, 0, 0, 1, 1, 1, 1]])


def test_sharded_job_needs_the_gpu_library():
    """No CPU fallback: a ShardedJob cannot be built without a CUDA device."""
    import torch

    from harmonica_b200 import distributed as hbd

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    coords, prisms, density = config1(10, 10, seed=1)
    with pytest.raises(Exception):
        hbd.ShardedJob("prism_gravity", coords, dict(prisms=prisms, density=density), "g_z").upload()


def test_shard_bounds_cover_everything():
    from harmonica_b200.distributed import shard_bounds

    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            edges = [shard_bounds(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges[:-1], edges[1:]))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
