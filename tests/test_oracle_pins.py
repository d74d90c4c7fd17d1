"""
Pins of the CPU oracle (oracle/choclo_port.c) against every known-answer the
reference's own tests and docs hold for the hot path, plus
formula-independent checks. CPU only.
"""

import numpy as np
import numpy.testing as npt
import pytest

import oracle as O
from _common import GRAVITY_FIELDS, golden, max_rel

G = 6.6743e-11


def test_point_potential_golden_csv():
    """reference test/test_point_gravity.py:144-154 + test/data/sample_point_gravity.csv"""
    g = golden("point_potential_csv")
    got = O.point_gravity((g["easting"], g["northing"], g["upward"]), g["point"], g["mass"], "potential")
    npt.assert_allclose(got, g["potential"])  # default rtol=1e-7, like the reference


def test_prism_gz_doctests():
    """reference src/harmonica/_forward/prisms/gravity.py:170-193 (5 decimals)"""
    coords = ([-40, 0, 40], [0, 0, 0], [30, 30, 30])
    gz = O.prism_gravity(coords, [-34, 5, -18, 14, -345, -146], 2670, "g_z")
    assert "({:.5f}, {:.5f}, {:.5f})".format(*gz) == "(0.06552, 0.06629, 0.06174)"
    prisms = [[-134, -5, -45, 45, -200, -50], [5, 134, -45, 45, -180, -30]]
    gz = O.prism_gravity(coords, prisms, [-300, 300], "g_z")
    assert "({:.5f}, {:.5f}, {:.5f})".format(*gz) == "(-0.05380, 0.02908, 0.11237)"


def test_prism_against_infinite_slab():
    """reference test/test_prism.py:269-297: observer ON the top face, sizes 1e3..1e9 m"""
    height, thickness, density = 1.5, 10.5, 2670
    sizes = np.logspace(3, 9, 7)
    results = np.array([
        O.prism_gravity((0, 0, height), [-s / 2, s / 2, -s / 2, s / 2, height - thickness, height],
                        density, "g_z") for s in sizes]).ravel()  # fmt: skip
    analytical = 1e5 * 2 * np.pi * G * density * thickness  # _gravity_corrections.py:16-78
    errors = abs(analytical - results)
    assert (errors[1:] < errors[:-1]).all()
    npt.assert_allclose(analytical, results[-1])


def test_prism_laplace():
    """reference test/test_prism.py:247-266"""
    e, n = np.meshgrid(np.linspace(-10e3, 10e3, 10), np.linspace(-10e3, 10e3, 10))
    coords = (e, n, np.full_like(e, 300.0))
    prisms = [[1e3, 7e3, -5e3, 2e3, -1e3, -500], [-4e3, 1e3, 4e3, 10e3, -2e3, 200]]
    rho = [2670.0, 2900.0]
    d = {f: O.prism_gravity(coords, prisms, rho, f) for f in ("g_ee", "g_nn", "g_zz")}
    npt.assert_allclose(d["g_ee"] + d["g_nn"], -d["g_zz"])


def test_prism_null_and_inverted():
    """reference test/test_prism.py:115-158: null prisms contribute nothing; an inverted
    prism (checks disabled) gives minus the potential"""
    coords = ([0.0, 30.0], [10.0, -20.0], [50.0, 60.0])
    p = np.array([[-100, 100, -100, 100, -200, -100.0]])
    base = O.prism_gravity(coords, p, [1000.0], "potential")
    both = O.prism_gravity(coords, np.vstack([p, [[5, 5, -1, 1, -3, -2.0]], [[0, 9, 0, 9, -9, -8.0]]]),
                           [1000.0, 2000.0, 0.0], "potential")
    npt.assert_array_equal(base, both)
    inv = O.prism_gravity(coords, [[100, -100, -100, 100, -200, -100.0]], [1000.0], "potential")
    npt.assert_allclose(inv, -base)


@pytest.mark.parametrize("field,dfield,axis", [("g_e", "potential", 0), ("g_n", "potential", 1),
                                                ("g_z", "potential", 2), ("g_ee", "g_e", 0),
                                                ("g_en", "g_e", 1), ("g_ez", "g_e", 2),
                                                ("g_nn", "g_n", 1), ("g_nz", "g_n", 2),
                                                ("g_zz", "g_z", 2)])  # fmt: skip
def test_prism_finite_differences(field, dfield, axis):
    """every kernel is the derivative of the one below it (SURVEY 8c mitigation)"""
    rng = np.random.default_rng(3)
    prisms = [[-300, 250, -180, 420, -900, -150.0]]
    rho = [2670.0]
    c = np.stack([rng.uniform(-2e3, 2e3, 30), rng.uniform(-2e3, 2e3, 30), rng.uniform(50, 800, 30)])
    h = 0.05
    cp, cm = c.copy(), c.copy()
    cp[axis] += h
    cm[axis] -= h
    fd = (O.prism_gravity(tuple(cp), prisms, rho, dfield) - O.prism_gravity(tuple(cm), prisms, rho, dfield)) / (2 * h)
    unit = {"potential": 1.0, "g_e": 1e-5, "g_n": 1e-5, "g_z": 1e-5}[dfield]
    out_unit = 1e5 if field in ("g_e", "g_n", "g_z") else 1e9
    fd = fd * unit * out_unit
    # harmonica's z points down: d/dz_down = -d/du; g_z itself is already "down"
    if axis == 2:
        fd = -fd
    got = O.prism_gravity(tuple(c), prisms, rho, field)
    npt.assert_allclose(got, fd, rtol=2e-6, atol=1e-9 * np.max(np.abs(got)))


def test_prism_potential_against_quadrature():
    """formula-independent pin: potential == G rho * triple integral of 1/r (mpmath)"""
    mpmath = pytest.importorskip("mpmath")
    mpmath.mp.dps = 20
    w, e, s, n, b, t = -30.0, 50.0, -20.0, 40.0, -80.0, -10.0
    E, N, U = 120.0, -75.0, 60.0
    val = mpmath.quad(lambda x, y, z: 1 / mpmath.sqrt((x - E) ** 2 + (y - N) ** 2 + (z - U) ** 2),
                      [w, e], [s, n], [b, t])
    want = float(val) * G * 2670.0
    got = float(O.prism_gravity((E, N, U), [w, e, s, n, b, t], 2670.0, "potential"))
    npt.assert_allclose(got, want, rtol=1e-10)


def test_tensor_face_rule_is_outside_limit():
    """gravity.py:153-158: on a face normal to a diagonal component the outside limit is returned"""
    prism = [-30.0, 50.0, -20.0, 40.0, -80.0, -10.0]
    eps = 1e-6
    for field, on, off in (("g_ee", (50.0, 5.0, -40.0), (50.0 + eps, 5.0, -40.0)),
                           ("g_ee", (-30.0, 5.0, -40.0), (-30.0 - eps, 5.0, -40.0)),
                           ("g_nn", (7.0, 40.0, -40.0), (7.0, 40.0 + eps, -40.0)),
                           ("g_nn", (7.0, -20.0, -40.0), (7.0, -20.0 - eps, -40.0)),
                           ("g_zz", (7.0, 5.0, -10.0), (7.0, 5.0, -10.0 + eps)),
                           ("g_zz", (7.0, 5.0, -80.0), (7.0, 5.0, -80.0 - eps))):  # fmt: skip
        a = float(O.prism_gravity(on, prism, 2670.0, field))
        b_ = float(O.prism_gravity(off, prism, 2670.0, field))
        npt.assert_allclose(a, b_, rtol=1e-5)


def test_tensor_nan_on_singular_points_and_warning_predicate():
    """test/test_prism.py:380-502: vertices are singular for every tensor component"""
    prism = np.array([[-30.0, 50.0, -20.0, 40.0, -80.0, -10.0]])
    vertex = (50.0, 40.0, -10.0)
    for f in GRAVITY_FIELDS[4:]:
        assert np.isnan(O.prism_gravity(vertex, prism, 2670.0, f))
        assert O.any_singular(vertex, prism, f)
    for f in GRAVITY_FIELDS[:4]:
        assert np.isfinite(O.prism_gravity(vertex, prism, 2670.0, f))
    # mid-point of an edge parallel to upward: singular for ee, nn, en only
    edge = (50.0, 40.0, -45.0)
    sing = {f: bool(np.isnan(O.prism_gravity(edge, prism, 2670.0, f))) for f in GRAVITY_FIELDS[4:]}
    assert sing == {"g_ee": True, "g_nn": True, "g_zz": False, "g_en": True, "g_ez": False, "g_nz": False}


def test_magnetic_far_field_is_dipole():
    """prism -> dipole with moment M*V far away (SURVEY 8a K4 far-field check)"""
    prism = [-1.0, 1.0, -1.5, 1.5, -2.0, 2.0]
    M = (np.array([1.3]), np.array([-0.4]), np.array([2.2]))
    vol = 2 * 3 * 4
    r = np.array([800.0, -500.0, 300.0])
    b = np.array(O.prism_magnetic(tuple(r), prism, M, "b")).ravel()
    m = np.array([M[0][0], M[1][0], M[2][0]]) * vol
    rn = np.linalg.norm(r)
    dip = 1e-7 * (3 * r * (m @ r) / rn**5 - m / rn**3) * 1e9
    npt.assert_allclose(b, dip, rtol=1e-4)


def test_magnetic_poisson_relation():
    """B = (mu0/4pi)/(G rho) * (grad grad V) . M  ties magnetics to the pinned gravity tensor"""
    rng = np.random.default_rng(5)
    prism = [-30.0, 50.0, -20.0, 40.0, -80.0, -10.0]
    c = (rng.uniform(-300, 300, 20), rng.uniform(-300, 300, 20), rng.uniform(5, 200, 20))
    M = np.array([0.7, -1.1, 0.4])
    T = {f: O.prism_gravity_si(tuple(np.ascontiguousarray(x) for x in c), np.array([prism]), np.array([1.0]), f) / G
         for f in GRAVITY_FIELDS[4:]}
    be = M[0] * T["g_ee"] + M[1] * T["g_en"] + M[2] * T["g_ez"]
    bn = M[0] * T["g_en"] + M[1] * T["g_nn"] + M[2] * T["g_nz"]
    bu = M[0] * T["g_ez"] + M[1] * T["g_nz"] + M[2] * T["g_zz"]
    got = np.array(O.prism_magnetic(c, prism, tuple(np.array([m]) for m in M), "b"))
    npt.assert_allclose(got, np.stack([be, bn, bu]) * 1e-7 * 1e9, rtol=1e-9, atol=0)


def test_point_symmetry_and_laplace():
    """test/test_point_gravity.py:354-395"""
    rng = np.random.default_rng(7)
    c = (rng.uniform(-1e3, 1e3, 40), rng.uniform(-1e3, 1e3, 40), rng.uniform(10, 500, 40))
    pts = ([0.0, 50.0], [10.0, -30.0], [-100.0, -300.0])
    m = [1e8, 3e8]
    d = {f: O.point_gravity(c, pts, m, f) for f in ("g_ee", "g_nn", "g_zz", "g_en", "g_ne", "g_ez", "g_ze")}
    npt.assert_allclose(d["g_ee"] + d["g_nn"], -d["g_zz"], atol=1e-9 * np.max(np.abs(d["g_zz"])))
    npt.assert_array_equal(d["g_en"], d["g_ne"])
    npt.assert_array_equal(d["g_ez"], d["g_ze"])


def test_point_spherical_analytic():
    """test/test_point_gravity.py:656-768: same-radial configuration has a closed form"""
    radius_p, mass = 6.0e6, 1e12
    radius = np.array([6.3e6, 6.5e6, 7.0e6])
    lon, lat = np.full(3, 13.0), np.full(3, -42.0)
    pot = O.point_gravity((lon, lat, radius), ([13.0], [-42.0], [radius_p]), [mass], "potential", "spherical")
    gz = O.point_gravity((lon, lat, radius), ([13.0], [-42.0], [radius_p]), [mass], "g_z", "spherical")
    npt.assert_allclose(pot, G * mass / (radius - radius_p), rtol=1e-9)
    npt.assert_allclose(gz, 1e5 * G * mass / (radius - radius_p) ** 2, rtol=1e-8)


def test_point_zero_distance_raises():
    """the reference's jitted loop raises ZeroDivisionError on a coincident pair"""
    with pytest.raises(ZeroDivisionError):
        O.point_gravity(([0.0], [0.0], [0.0]), ([0.0], [0.0], [0.0]), [1.0], "potential")
    with pytest.raises(ZeroDivisionError):
        O.eqs_predict(([1.0], [2.0], [3.0]), ([1.0], [2.0], [3.0]), [1.0])


def test_dipole_closed_forms():
    """dipole formula pins: on the dipole axis B = mu0/4pi * 2 m / d^3 along m; in the equatorial
    plane B = -mu0/4pi * m / d^3 (the reference's tests only compare against choclo)"""
    m = 3.5e6
    b = np.array(O.dipole_magnetic(([0.0], [0.0], [200.0]), ([0.0], [0.0], [0.0]), ([0.0], [0.0], [m]), "b")).ravel()
    npt.assert_allclose(b, [0.0, 0.0, 1e-7 * 2 * m / 200.0**3 * 1e9], rtol=1e-14, atol=1e-20)
    b = np.array(O.dipole_magnetic(([150.0], [0.0], [0.0]), ([0.0], [0.0], [0.0]), ([0.0], [0.0], [m]), "b")).ravel()
    npt.assert_allclose(b, [0.0, 0.0, -1e-7 * m / 150.0**3 * 1e9], rtol=1e-14, atol=1e-20)
    with pytest.raises(ZeroDivisionError):
        O.dipole_magnetic(([1.0], [2.0], [3.0]), ([1.0], [2.0], [3.0]), ([1.0], [0.0], [0.0]), "b")


def _gauss_legendre_box(bounds, order=64):
    """Tensor Gauss-Legendre nodes and weights of a box (formula-independent volume integrals)."""
    nodes, weights = np.polynomial.legendre.leggauss(order)
    axes = []
    for lo, hi in bounds:
        axes.append((0.5 * (hi - lo) * nodes + 0.5 * (hi + lo), 0.5 * (hi - lo) * weights))
    x, y, z = np.meshgrid(axes[0][0], axes[1][0], axes[2][0], indexing="ij")
    w = axes[0][1][:, None, None] * axes[1][1][None, :, None] * axes[2][1][None, None, :]
    return x, y, z, w


def test_prism_magnetic_against_volume_quadrature():
    """formula-independent pin of the ABSOLUTE values, units and signs of prism_magnetic (the
    reference's tests only compare with choclo, which is absent): B = mu0 / 4 pi times the volume
    integral of the dipole field 3 (M.R) R / R^5 - M / R^3, in nT, outside the prism"""
    prism = [-30.0, 50.0, -20.0, 40.0, -80.0, -10.0]
    magnetization = (1.7, -0.6, 2.3)  # A/m
    x, y, z, w = _gauss_legendre_box([(prism[0], prism[1]), (prism[2], prism[3]), (prism[4], prism[5])])
    for obs in ((120.0, -75.0, 60.0), (-110.0, 95.0, -40.0), (10.0, 15.0, 70.0)):
        rx, ry, rz = obs[0] - x, obs[1] - y, obs[2] - z
        r2 = rx * rx + ry * ry + rz * rz
        r5 = r2 ** 2.5
        m_dot_r = magnetization[0] * rx + magnetization[1] * ry + magnetization[2] * rz
        want = [1e-7 * 1e9 * np.sum(w * (3 * m_dot_r * rc / r5 - mc / r2 ** 1.5))
                for rc, mc in zip((rx, ry, rz), magnetization)]  # fmt: skip
        m = tuple(np.array([c]) for c in magnetization)
        got = O.prism_magnetic(obs, [prism], m, "b")
        npt.assert_allclose([float(np.ravel(c)[0]) for c in got], want, rtol=1e-9)
        for k, name in enumerate(("b_e", "b_n", "b_u")):
            npt.assert_allclose(float(np.ravel(O.prism_magnetic(obs, [prism], m, name))[0]), want[k], rtol=1e-9)


def test_prism_gravity_fields_against_volume_quadrature():
    """the same pin for the accelerations and the tensor: units (mGal, Eotvos) and signs
    (g_z down, g_ez / g_nz with the z axis down) straight from Newton's integral"""
    prism = [-30.0, 50.0, -20.0, 40.0, -80.0, -10.0]
    rho = 2670.0
    x, y, z, w = _gauss_legendre_box([(prism[0], prism[1]), (prism[2], prism[3]), (prism[4], prism[5])])
    for obs in ((120.0, -75.0, 60.0), (10.0, 15.0, 70.0)):
        dx, dy, dz = x - obs[0], y - obs[1], z - obs[2]  # from the observer to the mass
        r2 = dx * dx + dy * dy + dz * dz
        r3, r5 = r2 ** 1.5, r2 ** 2.5
        acc = {"g_e": dx / r3, "g_n": dy / r3, "g_z": -dz / r3}  # z axis DOWN: g_z = -g_u
        for field, integrand in acc.items():
            want = G * rho * 1e5 * np.sum(w * integrand)
            npt.assert_allclose(float(O.prism_gravity(obs, [prism], rho, field)), want, rtol=1e-9)
        tensor = {
            "g_ee": (3 * dx * dx - r2) / r5, "g_nn": (3 * dy * dy - r2) / r5, "g_zz": (3 * dz * dz - r2) / r5,
            "g_en": 3 * dx * dy / r5, "g_ez": -3 * dx * dz / r5, "g_nz": -3 * dy * dz / r5,
        }  # fmt: skip
        for field, integrand in tensor.items():
            want = G * rho * 1e9 * np.sum(w * integrand)
            npt.assert_allclose(float(O.prism_gravity(obs, [prism], rho, field)), want, rtol=1e-8,
                                atol=1e-12 * abs(want) + 1e-9)  # fmt: skip


def test_dipole_magnetic_against_prism_limit():
    """dipole_magnetic pinned on prism_magnetic: a small cube with magnetization M is the dipole
    m = M * volume seen from far away (relative difference ~ (size / distance)^2)"""
    half = 0.05
    prism = [-half, half, -half, half, -half, half]
    magnetization = (1.7, -0.6, 2.3)
    volume = (2 * half) ** 3
    obs = (40.0, -25.0, 30.0)
    m = tuple(np.array([c]) for c in magnetization)
    want = [float(np.ravel(c)[0]) for c in O.prism_magnetic(obs, [prism], m, "b")]
    moments = tuple(np.array([c * volume]) for c in magnetization)
    got = O.dipole_magnetic(obs, (np.array([0.0]), np.array([0.0]), np.array([0.0])), moments, "b")
    npt.assert_allclose([float(np.ravel(c)[0]) for c in got], want, rtol=1e-5)


def _sphere_moments(g):
    """dipole moments equivalent to the two magnetised spheres of the ellipsoid golden file"""
    volume = 4.0 / 3.0 * np.pi * float(g["radius"]) ** 3
    h0 = g["inducing_field"] * 1e-9 / float(g["mu_0"])
    chi = float(g["susceptibility"])
    return {"b_remanent": volume * g["remanent_mag"], "b_induced": volume * chi * h0 / (1 + chi / 3)}


def test_dipole_against_the_references_ellipsoid_code():
    """REFERENCE-HELD pin of dipole_magnetic that does not go through choclo: the reference's own
    ellipsoid_magnetic (numpy + scipy; golden made by oracle/make_golden_ellipsoid.py from the
    unmodified module) for uniformly magnetised spheres, whose external field is exactly a
    dipole's (the reference compares the two at 5e-4, test/ellipsoids/test_magnetic.py:560-581).
    Absolute values, nT and the signs of all three components agree to 1.4e-10; the whole
    difference is the digits of mu_0 (scipy's CODATA-2022 value in the ellipsoid code against
    4 pi 1e-7 here): rescaled by that ratio the two agree to rounding."""
    g = golden("ellipsoid_sphere_magnetic")
    coords = tuple(np.ascontiguousarray(c) for c in g["coordinates"])
    centre = tuple(np.array([c]) for c in g["centre"])
    ratio = float(g["mu_0"]) / (4 * np.pi * 1e-7)
    assert abs(ratio - 1) < 2e-10
    for key, moment in _sphere_moments(g).items():
        got = np.array(O.dipole_magnetic(coords, centre, tuple(np.array([m]) for m in moment), "b"))
        assert max_rel(got, g[key]) < 2e-10
        assert max_rel(got * ratio, g[key]) < 1e-13
        for k, field in enumerate(("b_e", "b_n", "b_u")):
            one = O.dipole_magnetic(coords, centre, tuple(np.array([m]) for m in moment), field)
            assert max_rel(np.asarray(one) * ratio, g[key][k]) < 1e-13


def test_point_accelerations_against_the_references_ellipsoid_code():
    """signs (g_z downward) and mGal scaling of point_gravity's accelerations against the
    reference's own ellipsoid_gravity (numpy + scipy, independent of the point kernels) for a
    homogeneous sphere, which is exactly a point mass outside
    (tests/golden/ellipsoid_sphere_gravity.npz, oracle/make_golden_ellipsoid.py). G itself is the
    constant both sides take from choclo and is not pinned by this."""
    g = golden("ellipsoid_sphere_gravity")
    coords = tuple(np.ascontiguousarray(c) for c in g["coordinates"])
    mass = 4.0 / 3.0 * np.pi * float(g["radius"]) ** 3 * float(g["density"])
    centre = tuple(np.array([c]) for c in g["centre"])
    for k, field in enumerate(("g_e", "g_n", "g_z")):
        got = O.point_gravity(coords, centre, np.array([mass]), field)
        assert max_rel(got, g["g_sphere"][k]) < 1e-14, field


def test_prism_potential_scale_against_the_references_tesseroid_code():
    """scale, sign and units of the prism POTENTIAL (no reference test holds an absolute value of
    it) against the reference's own tesseroid forward model (real numba code, no choclo) for a
    0.001 x 0.001 degree x 100 m tesseroid at the equator = a 111 x 111 x 100 m prism up to the
    curvature (1e-5) and the 0.1 % design accuracy of the tesseroid quadrature; g_z beside it
    (tests/golden/small_tesseroid.npz, oracle/make_golden_small_tesseroid.py)"""
    g = golden("small_tesseroid")
    R = float(g["mean_radius"])
    lon, lat, rad = np.radians(g["longitude"]), np.radians(g["latitude"]), g["radius"]
    # local Cartesian frame at (0, 0, R): easting, northing, upward
    coords = (rad * np.cos(lat) * np.sin(lon), rad * np.sin(lat), rad * np.cos(lat) * np.cos(lon) - R)
    w, e, s, n, bottom, top = g["tesseroid"][0]
    prism = np.array([[R * np.radians(w), R * np.radians(e), R * np.radians(s), R * np.radians(n),
                       bottom - R, top - R]])
    for field, bar in (("potential", 1e-3), ("g_z", 2e-3)):
        got = O.prism_gravity(coords, prism, g["density"], field)
        npt.assert_allclose(got, g[field], rtol=bar)
