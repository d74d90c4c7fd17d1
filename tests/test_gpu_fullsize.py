"""
BASELINE.json configs 3-5 at FULL size on one B200: the CUDA path is checked against the
oracle on an observer sample (the oracle finishes those in seconds) and through
size-independent properties on the full result (Laplace, linearity in the source strength,
NaN placement of the deterministic singular set, the reduce over source shards).
"""

import time
import warnings

import numpy as np
import numpy.testing as npt
import pytest

import oracle as O
from _common import TENSOR_FIELDS, TOL, config1, max_rel

pytestmark = pytest.mark.gpu


def test_config2_layer_observers_on_the_surface_all_fields(hb):
    """SURVEY 8d C2 parity variant at FULL size: the 500 x 500 layer observed ON its surface, at
    cell centres (on top faces) and at cell corners (on the shared vertical edges of four
    neighbouring prisms), all ten fields, NaN pattern of the tensor components included."""
    from _common import GRAVITY_FIELDS, layer_config2

    coords, east_c, north_c, bottom, top, density = layer_config2()
    rng = np.random.default_rng(22)
    half = (east_c[1] - east_c[0]) / 2
    cells = rng.permutation(500 * 500)
    k, j = np.divmod(cells, 500)
    ok = np.isfinite(top[k, j]) & (k < 499) & (j < 499)
    k, j = k[ok][:240], j[ok][:240]
    centres = (east_c[j[:120]], north_c[k[:120]], top[k[:120], j[:120]])
    corners = (east_c[j[120:]] + half, north_c[k[120:]] + half, top[k[120:], j[120:]])
    obs = tuple(np.concatenate([a, b]) for a, b in zip(centres, corners))
    nan_seen = False
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for f in GRAVITY_FIELDS:
            got = hb.prism_layer_gravity(obs, east_c, north_c, bottom, top, density, f)
            want = O.prism_layer_gravity(obs, east_c, north_c, bottom, top, density, f)
            assert max_rel(got, want) <= TOL, f  # compares the NaN patterns as well
            nan_seen |= bool(np.isnan(want).any())
            assert np.isfinite(want[:120]).all(), f  # face centres are never singular
    assert nan_seen  # corners sit on vertical edges: g_ee / g_nn / g_en are NaN there


def test_config3_full_tensor_1m_x_1m(hb):
    """prism_gravity, six tensor components fused, 1M prisms x 1M observers"""
    coords, prisms, density = config1(1_000_000, 1_000_000, seed=3, scale=10.0)
    t0 = time.perf_counter()
    ten = hb.prism_gravity(coords, prisms, density, TENSOR_FIELDS, disable_checks=True)
    dt = time.perf_counter() - t0
    print(f"config 3: 1e12 pairs x 6 components in {dt:.1f} s end to end = {1e12 / dt:.3e} pair/s")
    ten = np.stack(ten)
    assert ten.shape == (6, 1_000_000) and np.isfinite(ten).all()
    # Laplace on the full grid: g_ee + g_nn + g_zz = 0 outside the sources
    lap = ten[0] + ten[1] + ten[2]
    assert np.max(np.abs(lap)) <= 1e-9 * np.max(np.abs(ten[:3]))
    # oracle on a sample of observers (1M prisms x 48 observers x 6 fields)
    idx = np.random.default_rng(0).choice(1_000_000, 48, replace=False)
    sub = tuple(c[idx].copy() for c in coords)
    for k, f in enumerate(TENSOR_FIELDS):
        want = O.prism_gravity(sub, prisms, density, f)
        assert np.max(np.abs(ten[k, idx] - want)) <= TOL * np.max(np.abs(ten[k])), f


def test_config4_magnetic_200k_x_1m_with_singular_set(hb):
    """prism_magnetic b, 200k prisms x 1M observers + observers on vertices/edges/faces (C4)"""
    rng = np.random.default_rng(4)
    coords, prisms, _ = config1(200_000, 1_000_000, seed=4, scale=4.0)
    mag = tuple(rng.normal(size=200_000) for _ in range(3))
    # deterministic singular set: 8 vertices + 12 edge mid-points + 6 face centres of 160 prisms
    pts = []
    for w, e, s, n, b, t in prisms[:160]:
        xm, ym, zm = (w + e) / 2, (s + n) / 2, (b + t) / 2
        pts += [(x, y, z) for x in (w, e) for y in (s, n) for z in (b, t)]
        pts += [(xm, y, z) for y in (s, n) for z in (b, t)]
        pts += [(x, ym, z) for x in (w, e) for z in (b, t)]
        pts += [(x, y, zm) for x in (w, e) for y in (s, n)]
        pts += [(x, ym, zm) for x in (w, e)] + [(xm, y, zm) for y in (s, n)] + [(xm, ym, z) for z in (b, t)]
    sing = np.array(pts)
    assert sing.shape == (4160, 3)
    coords = tuple(np.concatenate([c, sing[:, k]]) for k, c in enumerate(coords))
    b = np.stack(hb.prism_magnetic(coords, prisms, mag, "b", disable_checks=True))
    assert b.shape == (3, 1_004_160)
    assert np.isfinite(b[:, :1_000_000]).all()
    # NaN exactly on the 20 edge/vertex points of each of the 160 prisms, finite on the face centres
    nan_rows = np.isnan(b[:, 1_000_000:]).all(axis=0).reshape(160, 26)
    assert nan_rows[:, :20].all() and not nan_rows[:, 20:].any()
    idx = np.concatenate([np.random.default_rng(1).choice(1_000_000, 64, replace=False),
                          1_000_000 + np.arange(0, 4160, 13)])
    sub = tuple(c[idx].copy() for c in coords)
    want = np.stack(O.prism_magnetic(sub, prisms, mag, "b"))
    scale = np.nanmax(np.abs(b), axis=1, keepdims=True)
    got = b[:, idx]
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.nanmax(np.abs(got - want) / scale) <= TOL
    # linearity in the magnetization on a slice of the observers
    part = tuple(c[:50_000] for c in coords)
    twice = np.stack(hb.prism_magnetic(part, prisms, tuple(2 * m for m in mag), "b", disable_checks=True))
    # (the 50k-observer call splits the source list into different chunks: rounding-level change)
    npt.assert_allclose(twice, 2 * b[:, :50_000], rtol=0, atol=1e-11 * np.nanmax(np.abs(b)))


def test_config5_eqs_predict_4m_x_4m(hb):
    """EquivalentSources.predict (sum coef/r) and point_gravity g_z, 4M sources x 4M observers (C5)"""
    rng = np.random.default_rng(5)
    n = 4_000_000
    side = 2000
    gx, gy = np.meshgrid(np.arange(side), np.arange(side))
    pe = (gx.ravel() + rng.uniform(-0.3, 0.3, n)) * 500.0
    pn = (gy.ravel() + rng.uniform(-0.3, 0.3, n)) * 500.0
    pu = np.full(n, -3000.0)
    coefs = rng.normal(size=n)
    coords = ((gx.ravel() + rng.uniform(-0.4, 0.4, n)) * 500.0,
              (gy.ravel() + rng.uniform(-0.4, 0.4, n)) * 500.0, rng.uniform(0, 500.0, n))
    t0 = time.perf_counter()
    pred = hb.eqs_predict(coords, (pe, pn, pu), coefs)
    dt = time.perf_counter() - t0
    print(f"config 5: 1.6e13 pairs in {dt:.1f} s end to end = {1.6e13 / dt:.3e} pair/s")
    assert pred.shape == (n,) and np.isfinite(pred).all()
    idx = rng.choice(n, 32, replace=False)
    sub = tuple(c[idx].copy() for c in coords)
    want = O.eqs_predict(sub, (pe, pn, pu), coefs)
    assert np.max(np.abs(pred[idx] - want)) <= TOL * np.max(np.abs(pred))
    # source-sharded evaluation by hand: two halves of the sources sum to the whole (linearity)
    part = tuple(c[:200_000] for c in coords)
    h = n // 2
    a = hb.eqs_predict(part, (pe[:h], pn[:h], pu[:h]), coefs[:h])
    b = hb.eqs_predict(part, (pe[h:], pn[h:], pu[h:]), coefs[h:])
    assert np.max(np.abs(a + b - pred[:200_000])) <= 1e-12 * np.max(np.abs(pred))
    # point_gravity g_z with the same geometry on a slice of the observers
    masses = np.abs(coefs) * 1e9
    gz = hb.point_gravity(part, (pe, pn, pu), masses, "g_z")
    want = O.point_gravity(tuple(c[:16] for c in part), (pe, pn, pu), masses, "g_z")
    assert np.max(np.abs(gz[:16] - want)) <= TOL * np.max(np.abs(gz))
