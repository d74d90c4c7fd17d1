"""
The product's per-pair math (harmonica_b200/csrc/hb200_math.cuh + hb200_fast.cuh)
compiled for the HOST (tests/harness) against the oracle. CPU only: this checks
the statements the CUDA kernels execute, not the kernels themselves (the GPU
parity tests do that).

  variant 0 (direct path) : bit-exact vs the oracle
  variant 1 (merged path) : within the north_star tolerance 1e-9 * max|field|
"""

import numpy as np
import numpy.testing as npt
import pytest

import oracle as O
from _common import GRAVITY_FIELDS, TOL, config1, golden, harness_prism, max_rel

G = 6.6743e-11
CM = 4 * np.pi * 1e-7 / 4 / np.pi


def _si(coords, prisms, density, field):
    return O.prism_gravity_si(tuple(np.ascontiguousarray(c, dtype=np.float64) for c in coords),
                              np.ascontiguousarray(prisms), np.ascontiguousarray(density), field)


@pytest.fixture(scope="module")
def case():
    coords, prisms, density = config1(300, 400, seed=11)
    prm = np.zeros((prisms.shape[0], 3))
    prm[:, 0] = G * density
    return coords, prisms, density, prm


@pytest.mark.parametrize("field", GRAVITY_FIELDS)
def test_direct_path_is_bit_exact(case, field):
    coords, prisms, density, prm = case
    out, _ = harness_prism(field, 0, coords, prisms, prm)
    npt.assert_array_equal(out[0], _si(coords, prisms, density, field))


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("field", GRAVITY_FIELDS)
def test_merged_path_within_tolerance(case, field, variant):
    coords, prisms, density, prm = case
    out, _ = harness_prism(field, variant, coords, prisms, prm)
    assert max_rel(out[0], _si(coords, prisms, density, field)) <= TOL


def test_fused_sets_match_single_fields(case):
    coords, prisms, density, prm = case
    for variant in (0, 1, 2):
        acc, _ = harness_prism("acc3", variant, coords, prisms, prm)
        ten, _ = harness_prism("tensor6", variant, coords, prisms, prm)
        for k, f in enumerate(GRAVITY_FIELDS[1:4]):
            assert max_rel(acc[k], _si(coords, prisms, density, f)) <= TOL
        for k, f in enumerate(GRAVITY_FIELDS[4:]):
            assert max_rel(ten[k], _si(coords, prisms, density, f)) <= TOL


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_magnetic(case, variant):
    coords, prisms, _, _ = case
    rng = np.random.default_rng(4)
    M = rng.normal(size=(prisms.shape[0], 3))
    want = np.array(O.prism_magnetic(coords, prisms, (M[:, 0], M[:, 1], M[:, 2]), "b"))
    got, _ = harness_prism("b", variant, coords, prisms, M)
    for k in range(3):
        assert max_rel(got[k] * CM * 1e9, want[k]) <= TOL
        single, _ = harness_prism(("b_e", "b_n", "b_u")[k], variant, coords, prisms, M)
        assert max_rel(single[0] * CM * 1e9, want[k]) <= TOL


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_observers_inside_below_and_around(variant):
    rng = np.random.default_rng(12)
    coords, prisms, density = config1(200, 300, seed=12)
    coords = (coords[0], coords[1], rng.uniform(-12e3, 2e3, 300))
    for t in range(40):  # strictly inside prism t
        coords[0][t] = 0.3 * prisms[t, 0] + 0.7 * prisms[t, 1]
        coords[1][t] = 0.6 * prisms[t, 2] + 0.4 * prisms[t, 3]
        coords[2][t] = 0.8 * prisms[t, 4] + 0.2 * prisms[t, 5]
    prm = np.zeros((200, 3))
    prm[:, 0] = G * density
    for f in GRAVITY_FIELDS:
        out, _ = harness_prism(f, variant, coords, prisms, prm)
        assert max_rel(out[0], _si(coords, prisms, density, f)) <= TOL, f


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_singular_suite_matches_reference_rules(variant):
    """vertices / edges / faces / edge extensions: NaN pattern and values (golden = reference)"""
    g = golden("prism_singular_suite")
    coords = (g["easting"], g["northing"], g["upward"])
    prm = np.zeros((2, 3))
    prm[:, 0] = G * g["density"]
    scale = {"potential": 1.0, "g_e": 1e5, "g_n": 1e5, "g_z": -1e5, "g_ee": 1e9, "g_nn": 1e9,
             "g_zz": 1e9, "g_en": 1e9, "g_ez": -1e9, "g_nz": -1e9}
    for f in GRAVITY_FIELDS:
        out, flags = harness_prism(f, variant, coords, g["prisms"], prm)
        assert max_rel(out[0] * scale[f], g[f"two_{f}"]) <= TOL, f
        assert bool(flags & 1) == (f in GRAVITY_FIELDS[4:])
    ten, _ = harness_prism("tensor6", variant, coords, g["prisms"], prm)
    for k, f in enumerate(GRAVITY_FIELDS[4:]):
        assert max_rel(ten[k] * scale[f], g[f"two_{f}"]) <= TOL, f
    b, flags = harness_prism("b", variant, coords, g["prisms"], g["mag"].T.copy())
    assert flags & 1
    for k in range(3):
        assert max_rel(b[k] * CM * 1e9, g["two_b"][k]) <= TOL


def test_slab_limit_on_face_merged_path():
    """test/test_prism.py:269-297 through the product math (observer on the top face)"""
    height, thickness, density = 1.5, 10.5, 2670
    sizes = np.logspace(3, 9, 7)
    res = []
    for s in sizes:
        out, _ = harness_prism("g_z", 2, ([0.0], [0.0], [height]),
                               [[-s / 2, s / 2, -s / 2, s / 2, height - thickness, height]],
                               [[G * density, 0, 0]])
        res.append(out[0][0] * -1e5)
    res = np.array(res)
    analytical = 1e5 * 2 * np.pi * G * density * thickness
    errors = abs(analytical - res)
    assert (errors[1:] < errors[:-1]).all()
    npt.assert_allclose(analytical, res[-1])
    # slightly above the face: the merged path proper, same limit
    out, _ = harness_prism("g_z", 2, ([0.0], [0.0], [height + 1e-3]),
                           [[-5e8, 5e8, -5e8, 5e8, height - thickness, height]], [[G * density, 0, 0]])
    npt.assert_allclose(out[0][0] * -1e5, analytical, rtol=1e-6)


def _mp_truth(field, E, N, U, prism, rho):
    """50-digit evaluation of the 8-vertex closed form (generic position only)."""
    import mpmath as mp

    w, e, s, n, b, t = [mp.mpf(float(x)) for x in prism]
    E, N, U = mp.mpf(float(E)), mp.mpf(float(N)), mp.mpf(float(U))
    tot = mp.mpf(0)
    for i, x in enumerate((e - E, w - E)):
        for j, y in enumerate((n - N, s - N)):
            for k, z in enumerate((t - U, b - U)):
                r = mp.sqrt(x * x + y * y + z * z)
                if field == "g_z":
                    v = -(x * mp.log(y + r) + y * mp.log(x + r) - z * mp.atan(x * y / (z * r)))
                else:  # potential
                    v = (x * y * mp.log(z + r) + y * z * mp.log(x + r) + x * z * mp.log(y + r)
                         - x * x / 2 * mp.atan(z * y / (x * r)) - y * y / 2 * mp.atan(z * x / (y * r))
                         - z * z / 2 * mp.atan(x * y / (z * r)))
                tot += (-1) ** (i + j + k) * v
    return mp.mpf("6.6743e-11") * mp.mpf(float(rho)) * tot


def test_large_region_accuracy_against_high_precision():
    """
    Region +-500 km with sparse, distant prisms: the field is tiny compared with
    the per-vertex terms (e * log(...) ~ 1e7 m), so the REFERENCE's own rounding
    noise exceeds 1e-9 * max|field| here. The merged path must (a) stay within
    that noise of the oracle and (b) be at least as close to a 50-digit
    evaluation as the oracle is.
    """
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    coords, prisms, density = config1(400, 60, seed=3, scale=10.0)
    prm = np.zeros((400, 3))
    prm[:, 0] = G * density
    for f in ("g_z", "g_zz", "g_en", "potential"):
        out, _ = harness_prism(f, 2, coords, prisms, prm)
        bound = {"g_zz": TOL, "g_en": TOL, "g_z": 5e-8, "potential": 1e-6}[f]
        assert max_rel(out[0], _si(coords, prisms, density, f)) <= bound, f
    for f in ("g_z", "potential"):
        out, _ = harness_prism(f, 2, coords, prisms, prm)
        ora = _si(coords, prisms, density, f)
        err_merged, err_oracle = [], []
        for i in range(4):
            t = sum(_mp_truth(f, coords[0][i], coords[1][i], coords[2][i], prisms[j], density[j])
                    for j in range(400))
            err_merged.append(float(abs(mp.mpf(float(out[0][i])) - t)))
            err_oracle.append(float(abs(mp.mpf(float(ora[i])) - t)))
        assert max(err_merged) <= max(err_oracle), (f, err_merged, err_oracle)


def test_dense_large_region_within_tolerance():
    """C3-like density of prisms near the observers: the 1e-9 bar holds against the oracle"""
    rng = np.random.default_rng(8)
    coords, prisms, density = config1(400, 60, seed=3, scale=10.0)
    # move the observers next to prisms (500 m above the top of a random prism)
    pick = rng.integers(0, 400, 60)
    coords = (0.5 * (prisms[pick, 0] + prisms[pick, 1]) + 300.0,
              0.5 * (prisms[pick, 2] + prisms[pick, 3]) - 200.0, prisms[pick, 5] + 500.0)
    prm = np.zeros((400, 3))
    prm[:, 0] = G * density
    # the potential is excluded: with coordinates ~5e5 m its per-vertex terms (e*n*log ~ 1e12)
    # put the reference's own rounding noise near 1e-8 * max|field| (see the test above)
    for f in GRAVITY_FIELDS[1:]:
        out, _ = harness_prism(f, 2, coords, prisms, prm)
        assert max_rel(out[0], _si(coords, prisms, density, f)) <= TOL, f


def test_xmath_sequences_accuracy():
    """hb200_xmath.cuh (host build): rcp <= 1 ulp, sqrt correctly rounded, log and atan2 <= 5e-16 abs"""
    import ctypes

    from _common import harness

    H = harness()
    dp = ctypes.POINTER(ctypes.c_double)
    H.hbt_xmath.argtypes = [ctypes.c_int, ctypes.c_int64, dp, dp, dp]

    def xm(op, a, b=None):
        a = np.ascontiguousarray(a, dtype=np.float64)
        b = a if b is None else np.ascontiguousarray(b, dtype=np.float64)
        out = np.empty_like(a)
        H.hbt_xmath(op, a.size, a.ctypes.data_as(dp), b.ctypes.data_as(dp), out.ctypes.data_as(dp))
        return out

    rng = np.random.default_rng(0)
    x = np.exp(rng.uniform(-60, 60, 400_000))
    assert np.max(np.abs(xm(0, x) - 1 / x) / np.spacing(1 / x)) <= 1.0
    npt.assert_array_equal(xm(1, x), np.sqrt(x))
    assert np.max(np.abs(xm(4, x) - np.sqrt(x)) / np.spacing(np.sqrt(x))) <= 1.0
    assert np.max(np.abs(xm(5, x) - 1 / np.sqrt(x)) / np.spacing(1 / np.sqrt(x))) <= 2.5  # the numpy reference has two roundings itself
    lg = xm(2, x)
    assert np.max(np.abs(lg - np.log(x)) / np.maximum(np.abs(np.log(x)), 1.0)) <= 4e-16
    near = 1 + rng.uniform(-2e-2, 2e-2, 400_000)
    assert np.max(np.abs(xm(2, near) - np.log(near))) <= 4e-18
    quarter = rng.uniform(-0.25, 0.25, 400_000)  # the table-free mid range (atanh form)
    want = np.log1p(quarter)
    assert np.max(np.abs(xm(6, quarter) - want) / np.abs(want)) <= 4.5e-16
    y, x2 = rng.normal(size=400_000) * 1e8, rng.normal(size=400_000) * 1e8
    assert np.max(np.abs(xm(3, y, x2) - np.arctan2(y, x2))) <= 5e-16
    small = rng.uniform(-1e-3, 1e-3, 400_000)  # far-field regime: tiny angles keep RELATIVE accuracy
    got = xm(3, small, np.ones_like(small))
    assert np.max(np.abs(got - np.arctan(small)) / np.abs(np.arctan(small))) <= 5e-16


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_near_axis_and_near_face_observers(variant):
    """observers ALMOST on the extension of an edge / almost in a face plane (offsets 1e-9 .. 1 m):
    the on-axis safe_log branch and the face rules must switch exactly where the reference's do"""
    prism = np.array([[-30.0, 50.0, -20.0, 40.0, -80.0, -10.0]])
    pts = []
    for off in (0.0, 1e-9, 1e-7, 1e-5, 1e-3, 1e-1, 1.0):
        pts += [(50.0 + 100.0, 40.0 + off, -10.0 + off), (-30.0 - 70.0, -20.0 - off, -80.0 + off),
                (50.0 + off, 40.0 + 250.0, -10.0 - off), (50.0 - off, 40.0 + off, -10.0 + 500.0),
                (10.0, 5.0, -10.0 + off), (50.0 + off, 5.0, -40.0), (10.0 + off, 40.0 + off, -10.0 + off)]
    a = np.array(pts)
    coords = (a[:, 0].copy(), a[:, 1].copy(), a[:, 2].copy())
    prm = np.array([[G * 2670.0, 0, 0]])
    for f in GRAVITY_FIELDS:
        out, _ = harness_prism(f, variant, coords, prism, prm)
        assert max_rel(out[0], _si(coords, prism, np.array([2670.0]), f)) <= TOL, f


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_fused_diagonal_components_inside_and_inverted_prisms(variant):
    """the fused tensor / magnetic sets derive k_uu from Laplace/Poisson (0 outside, -4 pi inside,
    sign flipped per inverted axis): check against the oracle with observers inside prisms,
    including prisms with inverted boundaries (disable_checks=True in the reference)"""
    rng = np.random.default_rng(14)
    coords, prisms, density = config1(60, 120, seed=14)
    coords = (coords[0], coords[1], rng.uniform(-12e3, 2e3, 120))
    for t in range(60):  # observer t strictly inside prism t
        coords[0][t] = 0.3 * prisms[t, 0] + 0.7 * prisms[t, 1]
        coords[1][t] = 0.6 * prisms[t, 2] + 0.4 * prisms[t, 3]
        coords[2][t] = 0.8 * prisms[t, 4] + 0.2 * prisms[t, 5]
    inv = prisms.copy()
    inv[:20, [0, 1]] = inv[:20, [1, 0]]          # one inverted axis
    inv[10:30, [4, 5]] = inv[10:30, [5, 4]]      # some with two, some with another single one
    M = rng.normal(size=(60, 3))
    prm = np.zeros((60, 3))
    prm[:, 0] = G * density
    for p in (prisms, inv):
        ten, _ = harness_prism("tensor6", variant, coords, p, prm)
        for k, f in enumerate(GRAVITY_FIELDS[4:]):
            assert max_rel(ten[k], _si(coords, p, density, f)) <= TOL, f
        b, _ = harness_prism("b", variant, coords, p, M)
        want = np.array(O.prism_magnetic(coords, p, (M[:, 0], M[:, 1], M[:, 2]), "b"))
        for k in range(3):
            assert max_rel(b[k] * CM * 1e9, want[k]) <= TOL


@pytest.mark.parametrize("scale", [1e-24, 1e-12, 1e12, 1e24, 1e60])
def test_extreme_length_scales_fall_back_to_the_reference_formulation(scale):
    """the merged products would leave the float64 range at absurd length scales; such pairs use
    the direct path, so the result is whatever the reference gives (no spurious inf/nan)"""
    coords, prisms, density = config1(40, 50, seed=15)
    coords = tuple(c * scale for c in coords)
    prisms = prisms * scale
    prm = np.zeros((40, 3))
    prm[:, 0] = G * density
    for f in ("g_z", "g_zz", "g_en", "potential"):
        out, _ = harness_prism(f, 2, coords, prisms, prm)
        want = _si(coords, prisms, density, f)
        assert np.isfinite(out[0]).all() == np.isfinite(want).all()
        assert max_rel(out[0], want) <= 1e-8, f


def _face_plane_case(seed, n_prisms=40, n_obs=360):
    """Observers in the plane of exactly ONE face of at least one prism: on the face itself,
    next to it (in the plane, outside the face) and level with it far away; prisms share bounds
    so that many pairs have one zero shift (flat tops, aligned walls)."""
    rng = np.random.default_rng(seed)
    # integer-valued bounds so that coincidences are exact; tops on two levels, walls on a lattice
    w = rng.integers(-20, 20, n_prisms) * 50.0
    s = rng.integers(-20, 20, n_prisms) * 50.0
    prisms = np.stack([w, w + rng.integers(1, 6, n_prisms) * 50.0, s,
                       s + rng.integers(1, 6, n_prisms) * 50.0,
                       -rng.integers(2, 9, n_prisms) * 100.0,
                       rng.integers(0, 2, n_prisms) * 100.0], axis=1)
    obs = np.empty((n_obs, 3))
    for t in range(n_obs):
        p = prisms[t % n_prisms]
        axis, side = (t // n_prisms) % 3, (t // (3 * n_prisms)) % 2
        where = t % 3  # 0: on the face, 1: in its plane just outside, 2: in its plane far away
        c = np.array([rng.uniform(p[0], p[1]), rng.uniform(p[2], p[3]), rng.uniform(p[4], p[5])])
        if where == 1:
            c += rng.uniform(300, 400, 3) * rng.choice([-1, 1], 3)
        elif where == 2:
            c += rng.uniform(3e3, 2e4, 3) * rng.choice([-1, 1], 3)
        c[axis] = p[2 * axis + 1 - side]  # east / north / top (side 0) or west / south / bottom
        # keep the other two coordinates off every lattice value: ONE zero shift only
        for a in range(3):
            if a != axis:
                c[a] += 0.123 + 0.01 * rng.uniform()
        obs[t] = c
    return (obs[:, 0].copy(), obs[:, 1].copy(), obs[:, 2].copy()), prisms, rng


@pytest.mark.parametrize("variant", [1, 2])
def test_one_zero_shift_stays_on_the_merged_path_with_the_face_rules(variant):
    """Observers in the plane of a prism face (stations on flat tops, grids aligned with prism
    walls): every field set against the oracle (= the reference's rules: +4 pi on the east /
    north / top face for the face-normal diagonal component), through the merged path."""
    coords, prisms, rng = _face_plane_case(21)
    density = rng.uniform(1000, 3000, prisms.shape[0])
    prm = np.zeros((prisms.shape[0], 3))
    prm[:, 0] = G * density
    for f in GRAVITY_FIELDS:
        out, flags = harness_prism(f, variant, coords, prisms, prm)
        assert max_rel(out[0], _si(coords, prisms, density, f)) <= TOL, f
        assert not flags & 1
    ten, _ = harness_prism("tensor6", variant, coords, prisms, prm)
    acc, _ = harness_prism("acc3", variant, coords, prisms, prm)
    for k, f in enumerate(GRAVITY_FIELDS[4:]):
        assert max_rel(ten[k], _si(coords, prisms, density, f)) <= TOL, f
    for k, f in enumerate(GRAVITY_FIELDS[1:4]):
        assert max_rel(acc[k], _si(coords, prisms, density, f)) <= TOL, f
    M = rng.normal(size=(prisms.shape[0], 3))
    for rules in (3, 1, 0):
        want = np.array(O.prism_magnetic(coords, prisms, (M[:, 0], M[:, 1], M[:, 2]), "b", flags=rules))
        got, _ = harness_prism("b", variant, coords, prisms, M, rules=rules)
        for k in range(3):
            assert max_rel(got[k] * CM * 1e9, want[k]) <= TOL, (rules, k)
            single, _ = harness_prism(("b_e", "b_n", "b_u")[k], variant, coords, prisms, M, rules=rules)
            assert max_rel(single[0] * CM * 1e9, want[k]) <= TOL, (rules, k)


def test_face_plane_pairs_are_classified_for_the_merged_path():
    """The point of the rule: a flat-topped model observed on its surface does not fall back to
    the rule-exact path (the direct path would give the same values, ~6x slower on the GPU)."""
    coords, prisms, rng = _face_plane_case(22)
    prm = np.zeros((prisms.shape[0], 3))
    prm[:, 0] = G
    d, _ = harness_prism("tensor6", 0, coords, prisms, prm)
    m, _ = harness_prism("tensor6", 2, coords, prisms, prm)
    assert max_rel(m, d) <= TOL
    assert not np.array_equal(m, d)  # evaluated by different code (merged vs direct)


@pytest.mark.parametrize("variant", [1, 2])
def test_negative_zero_bounds(variant):
    """A bound of -0.0 against a coordinate of +0.0 gives a shift of -0.0: same values."""
    prisms = np.array([[-0.0, 100.0, -50.0, 50.0, -80.0, -0.0], [-100.0, -0.0, -0.0, 70.0, -60.0, 0.0]])
    coords = (np.array([0.0, 20.0, -30.0, 0.0]), np.array([10.0, 0.0, 0.0, 25.0]),
              np.array([0.0, 0.0, 5.0, -20.0]))
    prm = np.array([[G * 2000.0, 0, 0], [G * 1500.0, 0, 0]])
    for f in ("potential", "g_z", "g_e", "g_n", "tensor6", "b"):
        want, fw = harness_prism(f, 0, coords, prisms, prm if f != "b" else np.ones((2, 3)))
        got, fg = harness_prism(f, variant, coords, prisms, prm if f != "b" else np.ones((2, 3)))
        assert max_rel(got, want) <= TOL, f
        assert fw == fg
