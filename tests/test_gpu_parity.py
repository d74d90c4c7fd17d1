"""
GPU parity tests: the CUDA path (through the public API -> ctypes -> C ABI ->
kernels) against the CPU oracle and the committed golden fixtures.

Tolerance (north_star): max abs error <= 1e-9 * max|field| (TOL), NaN patterns
identical. Run with `pytest -m gpu` on a B200.
"""

import ctypes
import warnings

import numpy as np
import numpy.testing as npt
import pytest

import oracle as O
from _common import (GRAVITY_FIELDS, TENSOR_FIELDS, TOL, config1, golden, layer_config2, max_rel,
                     random_prisms)

pytestmark = pytest.mark.gpu
G = 6.6743e-11


@pytest.fixture(params=[2, 1, 0], ids=["xmath", "merged", "direct"])
def variant(request, hb):
    lib = hb._lib.load()
    lib.hb200_set_variant(request.param)
    yield request.param
    lib.hb200_set_variant(2)


def quiet(fn, *a, **k):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn(*a, **k)


# ------------------------------------------------------------ golden fixtures
@pytest.mark.parametrize("where", ["above", "any"])
@pytest.mark.parametrize("field", GRAVITY_FIELDS)
def test_golden_prism_gravity(hb, variant, field, where):
    g = golden("prism_gravity_random")
    coords = tuple(g[f"{where}_{c}"] for c in "enu")
    got = hb.prism_gravity(coords, g["prisms"], g["density"], field)
    assert max_rel(got, g[f"{where}_{field}"]) <= TOL


@pytest.mark.parametrize("field", GRAVITY_FIELDS)
def test_golden_singular_suite(hb, variant, field):
    """observers on vertices, edges, faces, edge extensions, inside (reference values)"""
    g = golden("prism_singular_suite")
    coords = (g["easting"], g["northing"], g["upward"])
    if field in TENSOR_FIELDS:
        with pytest.warns(UserWarning, match="Found observation point on singular point of a prism."):
            got = hb.prism_gravity(coords, g["prisms"], g["density"], field)
    else:
        with warnings.catch_warnings():
            warnings.simplefilter("error")
            got = hb.prism_gravity(coords, g["prisms"], g["density"], field)
    assert max_rel(got, g[f"two_{field}"]) <= TOL
    got = quiet(hb.prism_gravity, coords, g["prisms"][0], g["density"][0], field)
    assert max_rel(got, g[f"one_{field}"]) <= TOL


@pytest.mark.parametrize("field", ["b", "b_e", "b_n", "b_u"])
def test_golden_prism_magnetic(hb, variant, field):
    for name, key in (("prism_magnetic_random", ""), ("prism_singular_suite", "two_")):
        g = golden(name)
        coords = (g["easting"], g["northing"], g["upward"])
        got = hb.prism_magnetic(coords, g["prisms"], tuple(g["mag"]), field)
        if field == "b":
            assert isinstance(got, tuple) and len(got) == 3
        assert max_rel(np.array(got), g[key + field]) <= TOL


@pytest.mark.parametrize("field", GRAVITY_FIELDS + ("g_ne", "g_ze", "g_zn"))
def test_golden_point_gravity(hb, field):
    g = golden("point_gravity_random")
    coords = (g["easting"], g["northing"], g["upward"])
    got = hb.point_gravity(coords, tuple(g["points"]), g["masses"], field)
    assert max_rel(got, g[field]) <= TOL


@pytest.mark.parametrize("field", ["potential", "g_z"])
def test_golden_point_gravity_spherical(hb, field):
    g = golden("point_gravity_random")
    got = hb.point_gravity(tuple(g["sph_obs"]), tuple(g["sph_points"]), g["masses"], field,
                           coordinate_system="spherical")
    assert max_rel(got, g[f"sph_{field}"]) <= TOL


def test_golden_point_potential_csv(hb):
    """the reference's own golden vector, test/test_point_gravity.py:144-154"""
    g = golden("point_potential_csv")
    got = hb.point_gravity((g["easting"], g["northing"], g["upward"]), g["point"], g["mass"], "potential")
    npt.assert_allclose(got, g["potential"])


@pytest.mark.parametrize("thr", [("thr0", None), ("thr10", 10.0)])
@pytest.mark.parametrize("field", GRAVITY_FIELDS)
def test_golden_prism_layer(hb, variant, field, thr):
    g = golden("prism_layer")
    coords = (g["easting"], g["northing"], g["upward"])
    with pytest.warns(UserWarning, match="Found NaN values in 'density'"):
        got = hb.prism_layer_gravity(coords, g["east_c"], g["north_c"], g["bottom"], g["top"],
                                     g["density"], field, thr[1])
    assert max_rel(got, g[f"{thr[0]}_{field}"]) <= TOL


def test_golden_eqs(hb):
    g = golden("eqs_predict")
    coords = (g["easting"], g["northing"], g["upward"])
    assert max_rel(hb.eqs_predict(coords, tuple(g["points"]), g["coefs"]), g["predicted"]) <= TOL
    npt.assert_allclose(hb.eqs_jacobian(coords, tuple(g["points"])), g["jacobian"], rtol=1e-14)
    res = np.ones(80)
    hb.predict_numba_parallel(coords, tuple(g["points"]), g["coefs"], res)
    assert max_rel(res - 1.0, g["predicted"]) <= 10 * TOL
    eqs = hb.EquivalentSources.from_fitted(tuple(g["points"]), g["coefs"])
    assert max_rel(eqs.predict(coords), g["predicted"]) <= TOL
    eqs32 = hb.EquivalentSources.from_fitted(tuple(g["points"]), g["coefs"], dtype="float32")
    out32 = eqs32.predict(coords)
    assert out32.dtype == np.float32
    npt.assert_allclose(out32, g["predicted"], rtol=2e-4)


# ------------------------------------------------ reference doctests / known answers
def test_reference_doctests_and_slab(hb, variant):
    """gravity.py:170-193 and test/test_prism.py:269-297 through the CUDA path"""
    coords = ([-40, 0, 40], [0, 0, 0], [30, 30, 30])
    gz = hb.prism_gravity(coords, [-34, 5, -18, 14, -345, -146], 2670, field="g_z")
    assert "({:.5f}, {:.5f}, {:.5f})".format(*gz) == "(0.06552, 0.06629, 0.06174)"
    gz = hb.prism_gravity(coords, [[-134, -5, -45, 45, -200, -50], [5, 134, -45, 45, -180, -30]],
                          [-300, 300], field="g_z")
    assert "({:.5f}, {:.5f}, {:.5f})".format(*gz) == "(-0.05380, 0.02908, 0.11237)"
    height, thickness, density = 1.5, 10.5, 2670
    sizes = np.logspace(3, 9, 7)
    res = np.array([
        hb.prism_gravity((0, 0, height), [-s / 2, s / 2, -s / 2, s / 2, height - thickness, height],
                         density, field="g_z") for s in sizes]).ravel()  # fmt: skip
    analytical = 1e5 * 2 * np.pi * G * density * thickness
    errors = abs(analytical - res)
    assert (errors[1:] < errors[:-1]).all()
    npt.assert_allclose(analytical, res[-1])


# ------------------------------------------------------------ seeded vs oracle
@pytest.fixture(scope="module")
def medium():
    coords, prisms, density = config1(1500, 2500, seed=21)
    # a third of the observers anywhere (below / inside the model)
    rng = np.random.default_rng(22)
    coords[2][:800] = rng.uniform(-11e3, 0, 800)
    return coords, prisms, density


@pytest.mark.parametrize("field", GRAVITY_FIELDS)
def test_prism_gravity_vs_oracle(hb, variant, medium, field):
    coords, prisms, density = medium
    got = hb.prism_gravity(coords, prisms, density, field)
    assert max_rel(got, O.prism_gravity(coords, prisms, density, field)) <= TOL


def test_fused_multi_field(hb, variant, medium):
    coords, prisms, density = medium
    ten = hb.prism_gravity(coords, prisms, density, TENSOR_FIELDS)
    acc = hb.prism_gravity(coords, prisms, density, ("g_z", "g_e", "g_n"))
    for f, got in zip(TENSOR_FIELDS, ten):
        assert max_rel(got, O.prism_gravity(coords, prisms, density, f)) <= TOL
    for f, got in zip(("g_z", "g_e", "g_n"), acc):
        assert max_rel(got, O.prism_gravity(coords, prisms, density, f)) <= TOL
    mixed = hb.prism_gravity(coords, prisms, density, ("potential", "g_zz"))
    assert max_rel(mixed[1], O.prism_gravity(coords, prisms, density, "g_zz")) <= TOL
    everything = hb.prism_gravity(coords, prisms, density, GRAVITY_FIELDS)  # 10 fields: two calls
    assert len(everything) == 10
    for f, got in zip(GRAVITY_FIELDS, everything):
        assert max_rel(got, O.prism_gravity(coords, prisms, density, f)) <= TOL, f


def test_prism_magnetic_vs_oracle(hb, variant, medium):
    coords, prisms, _ = medium
    rng = np.random.default_rng(23)
    M = tuple(rng.normal(size=prisms.shape[0]) for _ in range(3))
    want = np.array(O.prism_magnetic(coords, prisms, M, "b"))
    assert max_rel(np.array(hb.prism_magnetic(coords, prisms, M, "b")), want) <= TOL
    for k, f in enumerate(("b_e", "b_n", "b_u")):
        assert max_rel(hb.prism_magnetic(coords, prisms, M, f), want[k]) <= TOL


def test_magnetic_rule_switches(hb):
    g = golden("prism_singular_suite")
    coords = (g["easting"], g["northing"], g["upward"])
    for rules in (0, 1, 2, 3):
        got = np.array(hb.prism_magnetic(coords, g["prisms"], tuple(g["mag"]), "b", rules=rules))
        want = np.array(O.prism_magnetic(coords, g["prisms"], tuple(g["mag"]), "b", flags=rules))
        assert max_rel(got, want) <= TOL


def test_point_and_eqs_vs_oracle(hb):
    rng = np.random.default_rng(24)
    n_src, n_obs = 3000, 5001
    pts = (rng.uniform(-5e4, 5e4, n_src), rng.uniform(-5e4, 5e4, n_src), rng.uniform(-5e3, -1e3, n_src))
    m = rng.uniform(1e6, 1e9, n_src)
    coords = (rng.uniform(-5e4, 5e4, n_obs), rng.uniform(-5e4, 5e4, n_obs), rng.uniform(0, 500, n_obs))
    for f in GRAVITY_FIELDS:
        assert max_rel(hb.point_gravity(coords, pts, m, f), O.point_gravity(coords, pts, m, f)) <= TOL, f
    coefs = rng.normal(size=n_src)
    assert max_rel(hb.eqs_predict(coords, pts, coefs), O.eqs_predict(coords, pts, coefs)) <= TOL
    with pytest.raises(ZeroDivisionError):
        hb.point_gravity(([pts[0][5]], [pts[1][5]], [pts[2][5]]), pts, m, "potential")
    with pytest.raises(ZeroDivisionError):
        hb.eqs_predict(([pts[0][5]], [pts[1][5]], [pts[2][5]]), pts, coefs)


def test_layer_vs_oracle_and_flat(hb, variant):
    """layer.py accessor == oracle layer loop; == prism_gravity(_to_prisms()) (test_prism_layer.py:240-261)"""
    coords, east_c, north_c, bottom, top, density = layer_config2(n=40, seed=5)
    sub = tuple(c[::7].copy() for c in coords)
    layer = hb.PrismLayer((east_c, north_c), np.zeros((40, 40)), 0.0, properties={"density": density})
    layer.top, layer.bottom = top, bottom
    for f in GRAVITY_FIELDS:
        got = quiet(layer.prism_layer.gravity, sub, f)
        want = O.prism_layer_gravity(sub, east_c, north_c, bottom, top, density, f)
        assert max_rel(got, want) <= TOL, f
    # flat: prisms with NaN bounds or NaN/zero density removed by hand
    prisms = layer._to_prisms()
    rho = density.ravel()
    keep = ~(np.isnan(prisms).any(axis=1) | np.isnan(rho))
    flat = hb.prism_gravity(sub, prisms[keep], rho[keep], "g_z")
    npt.assert_allclose(quiet(layer.gravity, sub, "g_z"), flat, rtol=1e-9)
    # observers ON the surface: cell centres (top faces) and cell corners (shared vertical edges)
    ok = np.isfinite(top[:-1, :-1]) & np.isfinite(top[1:, :-1]) & np.isfinite(top[:-1, 1:]) & np.isfinite(top[1:, 1:])
    kk, jj = np.argwhere(ok)[np.argwhere(ok).shape[0] // 2]  # a cell whose three neighbours exist too
    e0, n0 = east_c[jj], north_c[kk]
    surf = (np.array([e0, e0 + 100.0]), np.array([n0, n0 + 100.0]), np.array([top[kk, jj], top[kk, jj]]))
    assert np.isfinite(surf[2]).all()  # the case must not vanish silently (seed-dependent NaN cells)
    for f in GRAVITY_FIELDS:
        got = quiet(layer.gravity, surf, f)
        want = O.prism_layer_gravity(surf, east_c, north_c, bottom, top, density, f)
        assert max_rel(got, want) <= TOL, f  # max_rel also compares the NaN patterns


# ----------------------------------------------------------------- edge cases
def test_shapes_dtypes_and_empty_inputs(hb):
    prism = [-34, 5, -18, 14, -345, -146]
    out = hb.prism_gravity((0, 0, 1000), prism, 2670, "g_z")
    assert out.shape == () and out.dtype == np.float64
    e, n = np.meshgrid(np.linspace(-50, 50, 6), np.linspace(-40, 40, 4))
    u = np.full_like(e, 30.0)
    out = hb.prism_gravity((e, n, u), prism, 2670, "g_z", dtype="float32")
    assert out.shape == (4, 6) and out.dtype == np.float32
    npt.assert_allclose(out, O.prism_gravity((e, n, u), prism, 2670, "g_z"), rtol=1e-6)
    # extra coordinate entries are ignored (coordinates[:3])
    out4 = hb.prism_gravity((e, n, u, np.zeros_like(e)), prism, 2670, "g_z")
    npt.assert_array_equal(out4, hb.prism_gravity((e, n, u), prism, 2670, "g_z"))
    # only null prisms -> exact zeros (gravity.py:452-486)
    z = hb.prism_gravity((e, n, u), [[0, 0, -1, 1, -2, -1], [-1, 1, -1, 1, -2, -1]], [5.0, 0.0], "g_z")
    assert (z == 0).all()
    zb = hb.prism_magnetic((e, n, u), [[-1, 1, -1, 1, -2, -1]], ([0.0], [0.0], [0.0]), "b")
    assert all((c == 0).all() for c in zb)
    # no observers
    assert hb.prism_gravity(([], [], []), prism, 2670, "g_z").shape == (0,)
    # integer inputs
    npt.assert_allclose(hb.prism_gravity(([0], [0], [30]), prism, 2670, "g_z"),
                        O.prism_gravity(([0.0], [0.0], [30.0]), prism, 2670, "g_z"), rtol=1e-12)
    # disable_checks: an inverted prism gives minus the potential (test_prism.py:139-158)
    a = hb.prism_gravity((e, n, u), [-100, 100, -100, 100, -200, -100], 1000, "potential")
    b = hb.prism_gravity((e, n, u), [100, -100, -100, 100, -200, -100], 1000, "potential",
                         disable_checks=True)
    npt.assert_allclose(b, -a, rtol=1e-9)


def test_singular_warning_rules(hb):
    """test/test_prism.py:380-502: which points warn for which component"""
    prism = [-30.0, 50.0, -20.0, 40.0, -80.0, -10.0]
    w, e, s, n, b, t = prism
    vertices = [(x, y, z) for x in (w, e) for y in (s, n) for z in (b, t)]
    for f in TENSOR_FIELDS:
        for v in vertices:
            with pytest.warns(UserWarning, match="Found observation point"):
                assert np.isnan(hb.prism_gravity(v, prism, 2670, f))
    singular_edges = {"g_ee": ("n", "u"), "g_nn": ("e", "u"), "g_zz": ("e", "n"), "g_en": ("u",),
                      "g_ez": ("n",), "g_nz": ("e",)}
    mid = {"e": (10.0, n, t), "n": (e, 10.0, b), "u": (w, s, -45.0)}
    for f, edges in singular_edges.items():
        for name, point in mid.items():
            with warnings.catch_warnings(record=True) as rec:
                warnings.simplefilter("always")
                val = hb.prism_gravity(point, prism, 2670, f)
            warned = any("singular point" in str(r.message) for r in rec)
            assert warned == (name in edges), (f, name)
            assert bool(np.isnan(val)) == (name in edges), (f, name)
    # disable_checks silences the warning but keeps the NaN
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        assert np.isnan(hb.prism_gravity(vertices[0], prism, 2670, "g_zz", disable_checks=True))
    # a NULL prism still triggers the warning (the scan runs before the discard), value stays finite
    with pytest.warns(UserWarning, match="Found observation point"):
        val = hb.prism_gravity(vertices[0], [prism, [-100, -90, 0, 10, -5, -1]], [0.0, 100.0], "g_zz")
    assert np.isfinite(val)


def test_progressbar_gives_identical_results(hb, medium):
    coords, prisms, density = medium
    a = hb.prism_gravity(coords, prisms[:200], density[:200], "g_z")
    b = hb.prism_gravity(coords, prisms[:200], density[:200], "g_z", progressbar=True)
    # the reference asserts allclose here too (test/test_prism.py:325-377). The value of a pair
    # does not depend on the batch (far-field sequences are chosen per lane), but the progress
    # bar's smaller observer batches split the source list into more chunks (grid.y), i.e. the
    # partial sums are associated differently
    npt.assert_allclose(a, b, rtol=0, atol=1e-11 * np.max(np.abs(a)))
    npt.assert_array_equal(a, hb.prism_gravity(coords, prisms[:200], density[:200], "g_z"))


def test_bit_reproducible_under_any_observer_batching(hb):
    """With the source chunking pinned (hb200_set_source_chunks) an observer's value is
    bit-identical whatever batch it is computed in: whole set, progress-bar chunks, a
    permutation, one observer at a time. (Far-field shortcuts are per-lane decisions; no
    warp votes.)"""
    lib = hb._lib.load()
    coords, prisms, density = config1(3000, 4099, seed=77)
    # near-field pairs too: some observers just above / next to prisms
    for t in range(64):
        coords[0][t] = prisms[t, 1] + 3.0
        coords[1][t] = 0.5 * (prisms[t, 2] + prisms[t, 3])
        coords[2][t] = prisms[t, 5] + 0.5
    try:
        for n_chunks in (1, 3):
            hb._lib.check(lib.hb200_set_source_chunks(n_chunks))
            for field in ("g_z", ("g_ee", "g_nn", "g_zz", "g_en", "g_ez", "g_nz"), "potential"):
                whole = np.stack(np.atleast_2d(hb.prism_gravity(coords, prisms, density, field)))
                bar = np.stack(np.atleast_2d(hb.prism_gravity(coords, prisms, density, field, progressbar=True)))
                npt.assert_array_equal(whole, bar)
                perm = np.random.default_rng(1).permutation(coords[0].size)
                shuffled = np.stack(np.atleast_2d(
                    hb.prism_gravity(tuple(c[perm] for c in coords), prisms, density, field)))
                npt.assert_array_equal(whole[:, perm], shuffled)
                for i in (0, 17, 63, 4098):
                    one = np.stack(np.atleast_2d(
                        hb.prism_gravity(tuple(c[i:i + 1] for c in coords), prisms, density, field)))
                    npt.assert_array_equal(whole[:, i:i + 1], one)
    finally:
        hb._lib.check(lib.hb200_set_source_chunks(0))


def test_few_observers_many_sources_uses_source_chunks(hb, variant):
    """small N: grid.y splits the source list; partials are reduced in a fixed order"""
    coords, prisms, density = config1(20000, 37, seed=31)
    for f in ("g_z", "g_en"):
        got = hb.prism_gravity(coords, prisms, density, f)
        assert max_rel(got, O.prism_gravity(coords, prisms, density, f)) <= TOL
        npt.assert_array_equal(got, hb.prism_gravity(coords, prisms, density, f))  # deterministic
    got = hb.prism_gravity(coords, prisms, density, "g_z", shard="sources")
    assert max_rel(got, O.prism_gravity(coords, prisms, density, "g_z")) <= TOL


# ---------------------------------------------- BASELINE sizes (properties + samples)
def test_config1_full_size(hb):
    """BASELINE config 1: 10k prisms x 10k observers, g_z, against the oracle in full"""
    coords, prisms, density = config1()
    got = hb.prism_gravity(coords, prisms, density, "g_z")
    want = O.prism_gravity(coords, prisms, density, "g_z")
    assert max_rel(got, want) <= TOL


def test_config2_layer_full_size_properties(hb):
    """BASELINE config 2 (500x500 layer, 250k observers): oracle on an observer sample,
    linearity in density, and Laplace's equation on the full grid"""
    coords, east_c, north_c, bottom, top, density = layer_config2()
    got = quiet(hb.prism_layer_gravity, coords, east_c, north_c, bottom, top, density, "g_z")
    assert got.shape == (250000,) and np.isfinite(got).all()
    idx = np.random.default_rng(0).choice(250000, 96, replace=False)
    sub = tuple(c[idx].copy() for c in coords)
    want = O.prism_layer_gravity(sub, east_c, north_c, bottom, top, density, "g_z")
    assert np.max(np.abs(got[idx] - want)) <= TOL * np.max(np.abs(got))
    twice = quiet(hb.prism_layer_gravity, coords, east_c, north_c, bottom, top, 2 * density, "g_z")
    npt.assert_allclose(twice, 2 * got, rtol=1e-12)
    small = tuple(c[::50].copy() for c in coords)
    d = {f: quiet(hb.prism_layer_gravity, small, east_c, north_c, bottom, top, density, f)
         for f in ("g_ee", "g_nn", "g_zz")}
    assert np.max(np.abs(d["g_ee"] + d["g_nn"] + d["g_zz"])) <= 1e-9 * np.max(np.abs(d["g_zz"]))


# ------------------------------------------------------- device-buffer entry points
def test_device_entry_points(hb):
    torch = pytest.importorskip("torch")
    lib = hb._lib.load()
    coords, prisms, density = config1(3000, 4096, seed=41)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    oe, on, ou, pr, rho = t(coords[0]), t(coords[1]), t(coords[2]), t(prisms), t(density)
    out = torch.empty((6, 4096), dtype=torch.float64, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    ws_bytes = lib.hb200_prism_ws_bytes(4096, 3000, 6)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    rc = lib.hb200_prism_gravity_dev(oe.data_ptr(), on.data_ptr(), ou.data_ptr(), 4096, pr.data_ptr(),
                                     rho.data_ptr(), 3000, 0x3F0, out.data_ptr(), flags.data_ptr(),
                                     ws.data_ptr(), ws_bytes, ctypes.c_void_p(stream))
    assert rc == 0, lib.hb200_last_error()
    torch.cuda.synchronize()
    res = out.cpu().numpy()
    for k, f in enumerate(TENSOR_FIELDS):
        assert max_rel(res[k], O.prism_gravity(coords, prisms, density, f)) <= TOL
    assert int(flags.item()) == 0


def test_fp64_peak_probe(hb):
    lib = hb._lib.load()
    flops, secs = ctypes.c_double(0), ctypes.c_double(0)
    assert lib.hb200_fp64_peak(2000, ctypes.byref(flops), ctypes.byref(secs)) == 0
    assert 5e12 < flops.value < 1e14


# ------------------------------------------- next rows of SURVEY 8f: dipoles, spherical EQS
@pytest.mark.parametrize("field", ["b", "b_e", "b_n", "b_u"])
def test_golden_dipole_magnetic(hb, field):
    g = golden("dipole_magnetic")
    coords = (g["easting"], g["northing"], g["upward"])
    got = hb.dipole_magnetic(coords, tuple(g["dipoles"]), tuple(g["moments"]), field)
    if field == "b":
        assert isinstance(got, tuple) and len(got) == 3
    assert max_rel(np.array(got), g[field]) <= TOL


def test_dipole_vs_oracle_and_zero_distance(hb):
    rng = np.random.default_rng(61)
    n_src, n_obs = 2500, 4099
    dip = (rng.uniform(-5e4, 5e4, n_src), rng.uniform(-5e4, 5e4, n_src), rng.uniform(-5e3, -1e2, n_src))
    mom = tuple(rng.normal(size=n_src) * 1e7 for _ in range(3))
    coords = (rng.uniform(-5e4, 5e4, n_obs), rng.uniform(-5e4, 5e4, n_obs), rng.uniform(0, 500, n_obs))
    want = np.array(O.dipole_magnetic(coords, dip, mom, "b"))
    assert max_rel(np.array(hb.dipole_magnetic(coords, dip, mom, "b")), want) <= TOL
    for k, f in enumerate(("b_e", "b_n", "b_u")):
        assert max_rel(hb.dipole_magnetic(coords, dip, mom, f), want[k]) <= TOL
    few = tuple(c[:11] for c in coords)  # few observers: chunked sources + reduce
    assert max_rel(np.array(hb.dipole_magnetic(few, dip, mom, "b")), want[:, :11]) <= TOL
    with pytest.raises(ZeroDivisionError):
        hb.dipole_magnetic(([dip[0][3]], [dip[1][3]], [dip[2][3]]), dip, mom, "b_u")


def test_golden_and_oracle_eqs_predict_spherical(hb):
    g = golden("eqs_predict_spherical")
    got = hb.eqs_predict(tuple(g["obs"]), tuple(g["points"]), g["coefs"], coordinate_system="spherical")
    assert max_rel(got, g["predicted"]) <= TOL
    eqs = hb.EquivalentSourcesSph.from_fitted(tuple(g["points"]), g["coefs"])
    assert max_rel(eqs.predict(tuple(g["obs"])), g["predicted"]) <= TOL
    res = np.zeros(80)
    hb.predict_numba_parallel(tuple(g["obs"]), tuple(g["points"]), g["coefs"], res,
                              hb._eqs.greens_func_spherical)
    assert max_rel(res, g["predicted"]) <= TOL
    rng = np.random.default_rng(62)
    n_src, n_obs = 3000, 5000
    pts = (rng.uniform(-40, 40, n_src), rng.uniform(-60, 60, n_src), rng.uniform(6.2e6, 6.3e6, n_src))
    obs = (rng.uniform(-45, 45, n_obs), rng.uniform(-65, 65, n_obs), rng.uniform(6.4e6, 6.5e6, n_obs))
    coefs = rng.normal(size=n_src)
    want = O.eqs_predict_spherical(obs, pts, coefs)
    assert max_rel(hb.eqs_predict(obs, pts, coefs, coordinate_system="spherical"), want) <= TOL


def test_equivalent_sources_fit_predict_round_trip(hb):
    """test/test_eq_sources_cartesian.py:157-183 (small data, Cartesian): fit on synthetic
    point-mass data and predict it back: GPU Jacobian + host least squares, GPU predict"""
    region = (-3e3, -1e3, 5e3, 7e3)

    def grid(shape, upward):
        e, n = np.meshgrid(np.linspace(region[0], region[1], shape[1]),
                           np.linspace(region[2], region[3], shape[0]))
        return e, n, np.full_like(e, float(upward))

    pts = grid((6, 6), -1e3)
    # checkerboard masses like verde.synthetic.CheckerBoard(amplitude=1e13, region=region)
    w_e, w_n = (region[1] - region[0]) / 2, (region[3] - region[2]) / 2
    masses = 1e13 * np.sin((2 * np.pi / w_e) * (pts[0] - region[0])) * np.cos(
        (2 * np.pi / w_n) * (pts[1] - region[2]))
    coords = grid((8, 8), 0)
    data = hb.point_gravity(coords, pts, masses, "g_z")
    eqs = hb.EquivalentSources(depth=500).fit(coords, data)
    # the interpolation should be perfect on the data points
    npt.assert_allclose(data, eqs.predict(coords), rtol=1e-5)
    npt.assert_allclose([c.ravel() for c in coords[:2]], eqs.points_[:2], rtol=1e-5)
    npt.assert_allclose(coords[2].ravel() - 500, eqs.points_[2], rtol=1e-5)
    assert eqs.depth_ == 500 and eqs.region_ == region
    up = grid((8, 8), 20)
    npt.assert_allclose(hb.point_gravity(up, pts, masses, "g_z"), eqs.predict(up), rtol=0.08)
    # default depth = 4.5 x mean first-neighbour distance; damped fit stays close to the data
    damped = hb.EquivalentSources(damping=1e-10).fit(coords, data)
    npt.assert_allclose(damped.depth_, 4.5 * (2e3 / 7), rtol=1e-12)
    npt.assert_allclose(damped.predict(coords), data, atol=1e-2 * np.max(np.abs(data)))
    # sources given explicitly (cartesian.py:270-275)
    fixed = hb.EquivalentSources(points=tuple(p.ravel() for p in pts)).fit(coords, data)
    assert fixed.depth_ is None and fixed.coefs_.shape == (36,)
    npt.assert_array_equal(fixed.points_[2], np.full(36, -1e3))
    # 36 sources with the 1/r kernel cannot reproduce 64 g_z values exactly: least-squares misfit
    assert np.max(np.abs(fixed.predict(coords) - data)) < 0.1 * np.max(np.abs(data))


def test_layer_vertex_reuse_is_bit_identical(hb):
    """tile mode 1 (default) against 2: neighbouring prisms of a layer column share two vertex distances
    (LayerCarry); same inputs to the same operations, so every field is bit-identical to the
    generic kernel — NaN / zero-density cells (broken chains), unequal bottoms, observers above,
    on top faces and on cell corners (exact-path pairs invalidate a lane's carry) included."""
    lib = hb._lib.load()
    coords, east_c, north_c, bottom, top, density = layer_config2(n=60, seed=7)
    rng = np.random.default_rng(3)
    bottom = bottom.copy()
    bottom[rng.integers(0, 60, 40), rng.integers(0, 60, 40)] -= 25.0  # unequal bottoms break the reuse
    ee, nn = np.meshgrid(east_c, north_c)
    ok = np.isfinite(top)
    on_top = (ee[ok][::7], nn[ok][::7], top[ok][::7])
    corners = (ee[ok][::11] + 100.0, nn[ok][::11] + 100.0, top[ok][::11])
    obs = tuple(np.concatenate([a[::5], b, c]) for a, b, c in zip(coords, on_top, corners))
    default = lib.hb200_get_tile_mode()
    try:
        for field in ("g_z", "potential", "g_e", "g_zz", "g_en"):
            results = []
            for mode in (2, 1):
                assert lib.hb200_set_tile_mode(mode) == 0
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    results.append(hb.prism_layer_gravity(obs, east_c, north_c, bottom, top, density, field))
            assert np.array_equal(results[0], results[1], equal_nan=True), field
    finally:
        lib.hb200_set_tile_mode(default)


def test_dipole_against_the_references_ellipsoid_code(hb):
    """the reference's own ellipsoid_magnetic for magnetised spheres (exactly dipoles outside;
    tests/golden/ellipsoid_sphere_magnetic.npz): a choclo-free pin of the absolute values, units
    and signs of dipole_magnetic; mu_0 digits account for 1.35e-10"""
    from test_oracle_pins import _sphere_moments

    g = golden("ellipsoid_sphere_magnetic")
    coords = tuple(np.ascontiguousarray(c) for c in g["coordinates"])
    centre = tuple(np.array([c]) for c in g["centre"])
    ratio = float(g["mu_0"]) / (4 * np.pi * 1e-7)
    for key, moment in _sphere_moments(g).items():
        got = np.array(hb.dipole_magnetic(coords, centre, tuple(np.array([m]) for m in moment), "b"))
        assert max_rel(got, g[key]) < TOL
        assert max_rel(got * ratio, g[key]) < 1e-12


def test_point_accelerations_against_the_references_ellipsoid_code(hb):
    """the reference's own ellipsoid_gravity for a homogeneous sphere (a point mass outside):
    signs and mGal scaling of point_gravity's accelerations (tests/golden/ellipsoid_sphere_gravity.npz)"""
    g = golden("ellipsoid_sphere_gravity")
    coords = tuple(np.ascontiguousarray(c) for c in g["coordinates"])
    mass = 4.0 / 3.0 * np.pi * float(g["radius"]) ** 3 * float(g["density"])
    centre = tuple(np.array([c]) for c in g["centre"])
    for k, field in enumerate(("g_e", "g_n", "g_z")):
        got = hb.point_gravity(coords, centre, np.array([mass]), field)
        assert max_rel(got, g["g_sphere"][k]) < 1e-12, field
