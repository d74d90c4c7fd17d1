"""
In-process multi-GPU sharding of the C library (hb200_init with several devices):
observer sharding (disjoint slices, no collective) and source sharding (peer copies to
device 0 + fixed-order reduce). Needs >= 2 B200s (run under `gpurun --gpus 2`).
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
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    hb.init(list(range(n)))
    yield hb
    hb.init([0])


def test_observer_and_source_sharding_match_single_gpu(hb_multi):
    hb = hb_multi
    coords, prisms, density = config1(6000, 40001, seed=51)
    multi_obs = hb.prism_gravity(coords, prisms, density, "g_z", shard="observers")
    multi_src = hb.prism_gravity(coords, prisms, density, "g_z", shard="sources")
    ten_obs = hb.prism_gravity(coords, prisms, density, TENSOR_FIELDS, shard="observers")
    ten_src = hb.prism_gravity(coords, prisms, density, TENSOR_FIELDS, shard="sources")
    n_dev = hb._lib.load().hb200_num_devices()
    hb.init([0])
    single = hb.prism_gravity(coords, prisms, density, "g_z")
    ten_single = hb.prism_gravity(coords, prisms, density, TENSOR_FIELDS)
    hb.init(list(range(n_dev)))
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
