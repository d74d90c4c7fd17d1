"""
In-process multi-GPU sharding of the C library (hb200_init with several devices):
observer sharding (disjoint slices, no collective) and source sharding (peer copies to
device 0 + fixed-order reduce). With >= 2 B200s (`gpurun --gpus 2`) the shards live on
different devices; on a 1-GPU box the same plans run as TWO shards on device 0
(`hb200_init([0, 0])`: two streams, two pools, the peer copy degenerates to a device copy),
so that run_host_job's sharding / gather / reduce code is always exercised.

Also: the torch.distributed path (harmonica_b200/distributed.py) as one local rank and, when
two GPUs are visible, as two NCCL ranks under torchrun.
"""

import numpy as np
import numpy.testing as npt
import pytest

import oracle as O
from _common import TENSOR_FIELDS, TOL, config1, max_rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hb_multi():
    import harmonica_b200 as hb

    lib = hb._lib.load()
    n = lib.hb200_device_count()
    devices = list(range(n)) if n >= 2 else [0, 0]
    hb.init(devices)
    hb._test_devices = devices
    yield hb
    hb.init([0])


def test_observer_and_source_sharding_match_single_gpu(hb_multi):
    hb = hb_multi
    coords, prisms, density = config1(6000, 40001, seed=51)
    multi_obs = hb.prism_gravity(coords, prisms, density, "g_z", shard="observers")
    multi_src = hb.prism_gravity(coords, prisms, density, "g_z", shard="sources")
    ten_obs = hb.prism_gravity(coords, prisms, density, TENSOR_FIELDS, shard="observers")
    ten_src = hb.prism_gravity(coords, prisms, density, TENSOR_FIELDS, shard="sources")
    assert hb._lib.load().hb200_num_devices() == len(hb._test_devices) >= 2
    hb.init([0])
    single = hb.prism_gravity(coords, prisms, density, "g_z")
    ten_single = hb.prism_gravity(coords, prisms, density, TENSOR_FIELDS)
    hb.init(hb._test_devices)
    # observer shards are computed by the same kernel on disjoint slices; the chunk
    # decomposition depends on the slice size, so agreement is to rounding, not bitwise
    assert max_rel(multi_obs, single) <= 1e-12
    assert max_rel(multi_src, single) <= 1e-12
    for a, b, c in zip(ten_obs, ten_src, ten_single):
        assert max_rel(a, c) <= 1e-12 and max_rel(b, c) <= 1e-12
    sub = tuple(c[:300] for c in coords)
    assert max_rel(multi_obs[:300], O.prism_gravity(sub, prisms, density, "g_z")) <= TOL


def test_other_entry_points_on_all_gpus(hb_multi):
    hb = hb_multi
    rng = np.random.default_rng(52)
    coords, prisms, _ = config1(3000, 30011, seed=53)
    M = tuple(rng.normal(size=3000) for _ in range(3))
    sub = tuple(c[:200] for c in coords)
    for shard in ("observers", "sources"):
        b = np.array(hb.prism_magnetic(coords, prisms, M, "b", shard=shard))
        assert max_rel(b[:, :200], np.array(O.prism_magnetic(sub, prisms, M, "b"))) <= TOL
    pts = (prisms[:, 0], prisms[:, 2], prisms[:, 4])
    coefs = rng.normal(size=3000)
    for shard in ("observers", "sources", "auto"):
        got = hb.eqs_predict(coords, pts, coefs, shard=shard)
        assert max_rel(got[:200], O.eqs_predict(sub, pts, coefs)) <= TOL
        g = hb.point_gravity(coords, pts, np.abs(coefs) * 1e9, "g_zz", shard=shard)
        assert max_rel(g[:200], O.point_gravity(sub, pts, np.abs(coefs) * 1e9, "g_zz")) <= TOL
    # deterministic: fixed reduce order
    a = hb.eqs_predict(coords, pts, coefs, shard="sources")
    npt.assert_array_equal(a, hb.eqs_predict(coords, pts, coefs, shard="sources"))


def test_layer_and_ragged_shards_on_all_devices(hb_multi):
    """The prism layer (replicated sources) and shard sizes that do not divide evenly."""
    hb = hb_multi
    from _common import layer_config2

    coords, ec, nc, bottom, top, density = layer_config2(n=60)
    pick = np.arange(0, coords[0].size, 3)[:1001]  # odd count: ragged observer shards
    sub = tuple(np.ascontiguousarray(c[pick]) for c in coords)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = hb.prism_layer_gravity(sub, ec, nc, bottom, top, density, "g_z")
        want = O.prism_layer_gravity(sub, ec, nc, bottom, top, density, "g_z")
    assert max_rel(got, want) <= TOL
    coords, prisms, density = config1(1001, 777, seed=54)
    for shard in ("observers", "sources"):
        got = hb.prism_gravity(coords, prisms, density, ("potential", "g_z"), shard=shard)
        for g, f in zip(got, ("potential", "g_z")):
            assert max_rel(g, O.prism_gravity(coords, prisms, density, f)) <= TOL


def test_sharded_job_single_rank(hb):
    """harmonica_b200.distributed without a process group: one rank, no collective."""
    from harmonica_b200 import distributed as hbd

    coords, prisms, density = config1(500, 2000, seed=55)
    got = hbd.prism_gravity(coords, prisms, density, ("g_z", "g_zz"))
    for g, f in zip(got, ("g_z", "g_zz")):
        assert max_rel(g, O.prism_gravity(coords, prisms, density, f)) <= TOL
    pts = (prisms[:, 0], prisms[:, 2], prisms[:, 4])
    assert max_rel(hbd.eqs_predict(coords, pts, density), O.eqs_predict(coords, pts, density)) <= TOL
    M = tuple(np.random.default_rng(1).normal(size=500) for _ in range(3))
    got = np.array(hbd.prism_magnetic(coords, prisms, M, "b"))
    assert max_rel(got, np.array(O.prism_magnetic(coords, prisms, M, "b"))) <= TOL
    job = hbd.ShardedJob("prism_gravity", coords, dict(prisms=prisms, density=density), "g_z")
    a = job.run()
    b = job.run()  # buffers are reused
    npt.assert_array_equal(a, b)


_TORCHRUN_WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
root = sys.argv[1]
sys.path[:0] = [root, os.path.join(root, "tests"), os.path.join(root, "oracle")]
import harmonica_b200 as hb
from harmonica_b200 import distributed as hbd
from _common import config1, layer_config2, max_rel
import oracle as O
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
hb.init([local])
coords, prisms, density = config1(3001, 4001, seed=56)   # ragged
res = hbd.prism_gravity(coords, prisms, density, ("g_z", "g_zz"))
pts = (prisms[:, 0], prisms[:, 2], prisms[:, 4])
eq = hbd.eqs_predict(coords, pts, density, shard="sources")
eq_all = hbd.eqs_predict(coords, pts, density, shard="sources", dst=None)
lc, ec, nc, bottom, top, rho = layer_config2(n=40)
import warnings
warnings.simplefilter("ignore")
lay = hbd.prism_layer_gravity(lc, ec, nc, bottom, top, rho, "g_z")
if dist.get_rank() == 0:
    sub = tuple(c[::7] for c in coords)
    for g, f in zip(res, ("g_z", "g_zz")):
        assert max_rel(g[::7], O.prism_gravity(sub, prisms, density, f)) <= 1e-9, f
    assert max_rel(eq[::7], O.eqs_predict(sub, pts, density)) <= 1e-9
    assert max_rel(lay, O.prism_layer_gravity(lc, ec, nc, bottom, top, rho, "g_z")) <= 1e-9
    print("TORCHRUN_OK")
else:
    assert res is None and eq is None and lay is None
assert eq_all is not None and eq_all.shape == (4001,)
dist.barrier()
dist.destroy_process_group()
"""


def test_torchrun_two_ranks_nccl(tmp_path):
    """Two NCCL ranks (one per GPU): observer shards gathered on rank 0, source shards reduced."""
    import os
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (NCCL refuses two ranks on one device)")
    from _common import ROOT

    script = tmp_path / "worker.py"
    script.write_text(_TORCHRUN_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", str(script), ROOT]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ))
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert "TORCHRUN_OK" in proc.stdout
