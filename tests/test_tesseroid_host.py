"""
Tesseroid forward model, CPU side (no GPU):

* the oracle (oracle/tesseroid_port.c) against the golden fixtures written by the reference's
  UNMODIFIED ``tesseroid_gravity`` (real numba, oracle/make_golden_tesseroid.py): bit-identical;
* oracle pins from the reference's own tests: spherical-shell closed form
  (test/test_tesseroid.py:664-770), discretisation monotonicity (:585-661), overflow errors
  (:317-336), distance / dimension closed forms (:377-423);
* the HOST BUILD of the product's pair function (csrc/hb200_tess.cuh, the statements the CUDA
  kernel executes) against the oracle: bit-identical values and leaf counts;
* the product's host-side checks (``harmonica_b200/_tesseroid.py``) against the reference's
  expectations (test/test_tesseroid.py:101-220, 339-462).
"""

import ctypes
import re

import numpy as np
import numpy.testing as npt
import pytest

import oracle as O
from _common import golden, harness

MEAN_RADIUS = 6371008.771415059  # boule.WGS84.mean_radius
G = 6.6743e-11
MODES = [("potential", False), ("potential", True), ("g_z", False), ("g_z", True)]


def _key(name, field, radial):
    return f"{name}_{field}_{'3d' if radial else '2d'}"


def _cases():
    g = golden("tesseroid")
    R = MEAN_RADIUS
    return g, {
        "random": (tuple(g["random_coords"]), g["random_tesseroids"], g["random_density"]),
        "doctest": ([0, 0, R], [-1.0, 1.0, -1.0, 1.0, R - 1000, R], 2670.0),
        "four": ([[-5.0, 0.0, 1.0], [-5.0, 0.0, 5.0], [R + 100] * 3],
                 [[-10.0, 0, -10.0, 0, R - 1e3, R], [-10.0, 0, 0, 10.0, R - 1e3, R],
                  [0, 10.0, -10.0, 0, R - 1e3, R], [0, 10.0, 0, 10.0, R - 1e3, R]],
                 1000.0 * np.ones(4)),
        "wrapped": ([0, 0, R + 1e3], [350, 10, -10, 10, R - 1e4, R], 1e3),
    }  # fmt: skip


def harness_tesseroid(coordinates, tesseroids, density, field, radial, density_upper=None):
    """Host build of hb200_tess.cuh summed like the kernel (SI, before sign / unit).
    density / density_upper: per radial quadrature node (equal for homogeneous tesseroids)."""
    H = harness()
    dp = ctypes.POINTER(ctypes.c_double)
    lon, lat, rad = (np.ascontiguousarray(np.atleast_1d(c), dtype=np.float64).ravel() for c in coordinates)
    tesseroids = np.ascontiguousarray(np.atleast_2d(tesseroids), dtype=np.float64)
    density = np.ascontiguousarray(np.atleast_1d(density), dtype=np.float64)
    upper = density if density_upper is None else np.ascontiguousarray(density_upper, dtype=np.float64)
    out = np.zeros(lon.size)
    counts = np.zeros((lon.size, tesseroids.shape[0]), dtype=np.int64)
    flags = ctypes.c_uint(0)
    H.hbt_tesseroid_loop(
        {"potential": 0, "g_z": 3}[field], ctypes.c_int64(lon.size), lon.ctypes.data_as(dp),
        lat.ctypes.data_as(dp), rad.ctypes.data_as(dp), ctypes.c_int64(tesseroids.shape[0]),
        tesseroids.ctypes.data_as(dp), density.ctypes.data_as(dp), upper.ctypes.data_as(dp), int(radial),
        out.ctypes.data_as(dp), counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
        ctypes.byref(flags),
    )  # fmt: skip
    return out, counts, flags.value


def reference_conditioning(coordinates, tesseroids, density, field, radial, trials=4):
    """
    How far the REFERENCE's own result moves (relative to max|field|) when its input coordinates
    change by one unit in the last place: its spherical distance goes through 1 - cos(psi), which
    loses (radius / distance)^2 * 1e-16 digits for nearby quadrature nodes, so two correct
    float64 evaluations with differently rounded sin / cos (glibc here, CUDA on the device) can
    only agree to this level. Used as the parity bar where it exceeds 1e-9.
    """
    coords = [np.atleast_1d(np.asarray(c, dtype=np.float64)) for c in coordinates]
    base = np.atleast_1d(O.tesseroid_gravity(coords, tesseroids, density, field, radial))
    worst = 0.0
    for k in range(trials):
        rng = np.random.default_rng(k)
        moved = [c * (1 + 2.2e-16 * rng.integers(-1, 2, c.shape)) for c in coords]
        out = np.atleast_1d(O.tesseroid_gravity(moved, tesseroids, density, field, radial))
        worst = max(worst, float(np.max(np.abs(out - base)) / np.max(np.abs(base))))
    return worst


def test_reference_conditioning_near_the_surface():
    """the 'four' case (points 100 m above 10-degree tesseroids): g_z moves by ~2e-9 of its
    maximum under a 1-ulp change of the inputs, the potential by ~1e-13"""
    _, cases = _cases()
    coords, tesseroids, density = cases["four"]
    assert reference_conditioning(coords, tesseroids, density, "g_z", False) > 2e-10
    assert reference_conditioning(coords, tesseroids, density, "potential", False) < 1e-11


# ------------------------------------------------------------------ oracle vs the reference
@pytest.mark.parametrize("field,radial", MODES)
@pytest.mark.parametrize("name", ["random", "doctest", "four", "wrapped"])
def test_oracle_is_bit_identical_to_the_reference(name, field, radial):
    g, cases = _cases()
    coords, tesseroids, density = cases[name]
    want = g[_key(name, field, radial)]
    if want.dtype.kind == "U":  # the reference raised (numba's division by zero)
        assert str(want) == "ZeroDivisionError"
        with pytest.raises(ZeroDivisionError):
            O.tesseroid_gravity(coords, tesseroids, density, field, radial)
        return
    got = O.tesseroid_gravity(coords, tesseroids, density, field, radial)
    assert np.array_equal(np.asarray(got).reshape(want.shape), want)


def test_oracle_leaves_match_the_reference():
    g = golden("tesseroid")
    tess = [-10.0, 10.0, -10.0, 10.0, 1.0, 10.0]
    for tag, radial in (("leaves_2d", False), ("leaves_3d", True)):
        setup = g[tag + "_setup"]
        got = O.adaptive_discretization(setup[:3], tess, setup[3], radial)
        assert np.array_equal(got, g[tag])
    # test/test_tesseroid.py:642-661: 2-D discretisation never splits the radial direction
    assert np.all(g["leaves_2d"][:, 4] == 1.0) and np.all(g["leaves_2d"][:, 5] == 10.0)


def test_doctest_value():
    """tesseroid_gravity.py:170-183: the value the reference's doctest computes."""
    g = golden("tesseroid")
    R = MEAN_RADIUS
    got = O.tesseroid_gravity([0, 0, R], [-1.0, 1.0, -1.0, 1.0, R - 1000, R], 2670.0, "g_z")
    assert float(got) == float(g["doctest_g_z_2d"])
    npt.assert_allclose(got, 112.567, atol=1e-3)  # within 0.6 % of the Bouguer slab 2 pi G rho h


# ------------------------------------------------------------------ oracle pins
def _shell(shape, thickness):
    longitude = np.linspace(0, 360, shape[0] + 1)
    latitude = np.linspace(-90, 90, shape[1] + 1)
    top = MEAN_RADIUS
    return np.array([[w, e, s, n, top - thickness, top]
                     for w, e in zip(longitude[:-1], longitude[1:])
                     for s, n in zip(latitude[:-1], latitude[1:])])  # fmt: skip


def _shell_analytical(top, bottom, density, radius):
    potential = 4 / 3 * np.pi * G * density * (top**3 - bottom**3) / radius
    return {"potential": potential, "g_z": 1e5 * potential / radius}


@pytest.mark.parametrize("field", ["potential", "g_z"])
@pytest.mark.parametrize("thickness", [10, 1e3, 1e5])
def test_spherical_shell_three_dim_adaptive_discret(field, thickness):
    """test/test_tesseroid.py:729-770"""
    radius = MEAN_RADIUS + 1e3
    tesseroids = _shell((6, 6), thickness)
    want = _shell_analytical(MEAN_RADIUS, MEAN_RADIUS - thickness, 1000, radius)[field]
    got = O.tesseroid_gravity([0, 0, radius], tesseroids, 1000 * np.ones(36), field, True)
    npt.assert_allclose(got, want, rtol=1e-3)


@pytest.mark.parametrize("field", ["potential", "g_z"])
def test_spherical_shell_two_dim_adaptive_discret(field):
    """test/test_tesseroid.py:683-726 on a thinned grid of points ON the shell"""
    lon, lat = np.meshgrid(np.arange(0, 351, 70.0), np.arange(-90, 91, 45.0))
    coords = (lon.ravel(), lat.ravel(), np.full(lon.size, MEAN_RADIUS))
    for thickness in (10, 1e3, 1e5):
        tesseroids = _shell((12, 6), thickness)
        want = _shell_analytical(MEAN_RADIUS, MEAN_RADIUS - thickness, 1000, MEAN_RADIUS)[field]
        got = O.tesseroid_gravity(coords, tesseroids, 1000 * np.ones(72), field)
        npt.assert_allclose(got, want, rtol=1e-3)


@pytest.mark.parametrize("radial", [True, False])
def test_adaptive_discretization_monotonic(radial):
    """test/test_tesseroid.py:585-639"""
    tess = [-10.0, 10.0, -10.0, 10.0, 1.0, 10.0]
    radii = [10.1 if radial else 10.0, 10.5, 12.0, 13.0, 15.0, 20.0, 30.0]
    counts = [O.adaptive_discretization([0.0, 0.0, r], tess, 10, radial).shape[0] for r in radii]
    assert all(a >= b for a, b in zip(counts, counts[1:]))
    counts = [O.adaptive_discretization([0.0, 0.0, 10.2], tess, ratio, radial).shape[0]
              for ratio in np.linspace(1, 10, 10)]  # fmt: skip
    assert all(a <= b for a, b in zip(counts, counts[1:]))


def test_overflow_errors():
    """test/test_tesseroid.py:317-336"""
    tess, point = [-10.0, 10.0, -10.0, 10.0, 0.5, 1.0], [0.0, 0.0, 1.0]
    with pytest.raises(OverflowError, match="Stack Overflow"):
        O.adaptive_discretization(point, tess, 10, stack_size=2)
    with pytest.raises(OverflowError, match="Exceeded maximum discretizations"):
        O.adaptive_discretization(point, tess, 10, max_small=2)


def test_distance_and_dimensions_closed_forms():
    """test/test_tesseroid.py:377-423"""
    L = O.lib()
    dp = ctypes.POINTER(ctypes.c_double)

    def distance(point, tess):
        p, t = np.array(point, dtype=float), np.array(tess, dtype=float)
        return L.hbo_distance_tesseroid_point(p.ctypes.data_as(dp), t.ctypes.data_as(dp))

    R = MEAN_RADIUS
    tess = [-1.0, 1.0, -1.0, 1.0, R - 0.65, R + 0.65]
    npt.assert_allclose(distance([0.0, 0.0, R], tess), 0.0)
    npt.assert_allclose(distance([0.0, 0.0, R + 1.0], tess), 1.0)
    npt.assert_allclose(distance([3.0, 0.0, R], tess), 2 * R * np.sin(0.5 * np.radians(3.0)))
    npt.assert_allclose(distance([0.0, 3.0, R], tess), 2 * R * np.sin(0.5 * np.radians(3.0)))
    t = np.array([-1.0, 1.0, -1.0, 1.0, 0.5, 1.5])
    dims = [ctypes.c_double() for _ in range(3)]
    L.hbo_tesseroid_dimensions(t.ctypes.data_as(dp), *[ctypes.byref(d) for d in dims])
    npt.assert_allclose([d.value for d in dims], [1.5 * np.radians(2.0), 1.5 * np.radians(2.0), 1.0])


# ------------------------------------------------------------------ product math (host build)
@pytest.mark.parametrize("field,radial", MODES)
def test_product_pair_function_is_bit_identical_to_the_oracle(field, radial):
    g, cases = _cases()
    for name in ("random", "four", "wrapped"):
        coords, tesseroids, density = cases[name]
        tesseroids = np.atleast_2d(np.asarray(tesseroids, dtype=float))
        if (tesseroids[:, 0] > tesseroids[:, 1]).any():
            tesseroids = O.longitude_continuity(tesseroids)
        density = np.atleast_1d(np.asarray(density, dtype=float))
        want, want_counts = O.tesseroid_gravity(coords, tesseroids, density, field, radial, return_counts=True)
        got, counts, flags = harness_tesseroid(coords, tesseroids, density, field, radial)
        if field == "g_z":
            got *= -1
            got *= 1e5
        assert flags == 0
        assert np.array_equal(counts, want_counts)
        assert np.array_equal(got, np.asarray(want).ravel())


def test_product_pair_function_random_models():
    rng = np.random.default_rng(77)
    R = MEAN_RADIUS
    for trial in range(4):
        n_tess, n_obs = 30, 40
        w, s = rng.uniform(-170, 160, n_tess), rng.uniform(-85, 75, n_tess)
        tesseroids = np.stack([w, w + rng.uniform(0.1, 10, n_tess), s, s + rng.uniform(0.1, 10, n_tess),
                               R - rng.uniform(1e3, 1e5, n_tess), R - rng.uniform(0, 500, n_tess)], 1)  # fmt: skip
        density = rng.uniform(-1000, 3000, n_tess)
        coords = (rng.uniform(-180, 180, n_obs), rng.uniform(-90, 90, n_obs),
                  R + rng.uniform(0, 10.0 ** rng.uniform(1, 6), n_obs))  # fmt: skip
        for field, radial in MODES:
            want, want_counts = O.tesseroid_gravity(coords, tesseroids, density, field, radial, return_counts=True)
            got, counts, flags = harness_tesseroid(coords, tesseroids, density, field, radial)
            if field == "g_z":
                got *= -1
                got *= 1e5
            assert flags == 0 and np.array_equal(counts, want_counts) and np.array_equal(got, want)


def test_product_pair_function_flags():
    """where the reference raises, the pair function reports a flag (and terminates)"""
    R = MEAN_RADIUS
    # a point on the top-face centre with 3-D discretisation: numba's ZeroDivisionError
    _, _, flags = harness_tesseroid([0, 0, R], [-1.0, 1.0, -1.0, 1.0, R - 1000, R], 2670.0, "g_z", True)
    assert flags & 2
    with pytest.raises(ZeroDivisionError):
        O.tesseroid_gravity([0, 0, R], [-1.0, 1.0, -1.0, 1.0, R - 1000, R], 2670.0, "g_z", True)
    # test/test_tesseroid.py:317-336: a stack of 2 / a leaf budget of 2 overflow
    H = harness()
    dp = ctypes.POINTER(ctypes.c_double)
    H.hbt_tesseroid_overflow.restype = ctypes.c_uint
    H.hbt_tesseroid_overflow.argtypes = [ctypes.c_int, dp, dp, ctypes.c_double]
    tess, point = np.array([-10.0, 10.0, -10.0, 10.0, 0.5, 1.0]), np.array([0.0, 0.0, 1.0])
    assert H.hbt_tesseroid_overflow(0, point.ctypes.data_as(dp), tess.ctypes.data_as(dp), 10.0) == 4
    assert H.hbt_tesseroid_overflow(1, point.ctypes.data_as(dp), tess.ctypes.data_as(dp), 10.0) == 8


# ------------------------------------------------------------------ product host checks
def test_check_tesseroids_valid_and_invalid():
    """test/test_tesseroid.py:134-220"""
    from harmonica_b200._tesseroid import _check_tesseroids

    w, e, s, n, bottom, top = -10, 10, -10, 10, 100, 200
    for tess in ([w, e, s, n, bottom, top], [w, w, s, n, bottom, top], [w, e, s, s, bottom, top],
                 [w, e, s, n, bottom, bottom], [350, 10, s, n, bottom, top], [-70, -60, s, n, bottom, top],
                 [-150, 150, s, n, bottom, top], [0, 360, s, n, bottom, top], [-180, 180, s, n, bottom, top]):  # fmt: skip
        _check_tesseroids(np.atleast_2d(np.array(tess, dtype=float)))
    bad = [
        ("The south boundary can't be greater than the north one", [w, e, n, s, bottom, top]),
        ("The latitudinal boundaries must be inside the [-90, 90] degrees interval", [w, e, s, -100, bottom, top]),
        ("The latitudinal boundaries must be inside the [-90, 90] degrees interval", [w, e, 100, n, bottom, top]),
        ("The bottom radius boundary can't be greater than the top one", [w, e, s, n, top, bottom]),
        ("The bottom and top radii should be positive or zero", [w, e, s, n, bottom, -1]),
        ("The bottom and top radii should be positive or zero", [w, e, s, n, -1, top]),
        ("The longitudinal boundaries must be inside the [-180, 360] degrees interval", [-200, e, s, n, bottom, top]),
        ("The longitudinal boundaries must be inside the [-180, 360] degrees interval", [w, 400, s, n, bottom, top]),
        ("The west boundary can't be greater than the east one", [30, 0, s, n, bottom, top]),
        ("The west boundary can't be greater than the east one", [-60, -70, s, n, bottom, top]),
        ("The west boundary can't be greater than the east one", [300, -150, s, n, bottom, top]),
        ("The west boundary can't be greater than the east one", [350, 340, s, n, bottom, top]),
        ("The difference between east and west boundaries cannot be greater than one turn around the globe",
         [-150, 300, s, n, bottom, top]),
    ]  # fmt: skip
    for msg, tess in bad:
        with pytest.raises(ValueError, match=re.escape(msg)):
            _check_tesseroids(np.atleast_2d(np.array(tess, dtype=float)))


def test_longitude_continuity_and_null_tesseroids():
    """test/test_tesseroid.py:339-441"""
    from harmonica_b200._tesseroid import _discard_null_tesseroids, _longitude_continuity

    for tess, want in (([-10, 10], (-10, 10)), ([-70, -60], (-70, -60)), ([350, 10], (-10, 10))):
        out = _longitude_continuity(np.atleast_2d(np.array(tess + [-10, 10, 1, 2], dtype=float)))
        assert (out[0, 0], out[0, 1]) == want
    top, bottom = MEAN_RADIUS, MEAN_RADIUS - 1e3
    tesseroids = np.array([
        [-10, -5, -10, -5, bottom, top], [-10, -5, -5, 0, bottom, top], [-10, -10, 0, 5, bottom, top],
        [-10, -5, 5, 5, bottom, top], [-5, 0, -10, -5, top, top], [-5, 0, -5, 0, bottom, top],
        [-5, -5, 5, 5, top, top], [-5, 0, 5, 10, bottom, top]])  # fmt: skip
    densities = np.array([2400, 0, 2500, 2600, 2700, 2800, 2900, 3000])
    tesseroids, densities = _discard_null_tesseroids(tesseroids, densities)
    npt.assert_allclose(tesseroids, [[-10, -5, -10, -5, bottom, top], [-5, 0, -5, 0, bottom, top],
                                     [-5, 0, 5, 10, bottom, top]])  # fmt: skip
    npt.assert_allclose(densities, [2400, 2800, 3000])


def test_points_inside_tesseroids_predicate():
    """test/test_tesseroid.py:247-314 through the pair predicate the device scan evaluates
    (host build) and the host-side pair listing"""
    from harmonica_b200._tesseroid import _conflicting_pairs

    tesseroid = np.atleast_2d(np.array([-10, 10, -10, 10, 100, 200], dtype=float))
    outside = np.array([[0, 0, 250], [20, 0, 150], [0, 20, 150], [0, 0, 200], [0, 0, 100], [-10, 0, 150],
                        [10, 0, 150], [0, -10, 150], [0, 10, 150]], dtype=float).T  # fmt: skip
    assert _conflicting_pairs(tuple(outside), tesseroid) == []
    for point in ([0, 0, 150], [360, 0, 150]):
        assert _conflicting_pairs(tuple(np.atleast_2d(np.array(point, dtype=float)).T), tesseroid) == [(0, 0)]
    phased = np.atleast_2d(np.array([260, 280, -10, 10, 100, 200], dtype=float))
    assert _conflicting_pairs(tuple(np.atleast_2d(np.array([-90.0, 0, 150])).T), phased) == [(0, 0)]
    tesseroids = np.array([[-10, 10, -10, 10, 100, 200], [20, 30, 20, 30, 400, 500],
                           [-50, -40, -30, -20, 100, 500]], dtype=float)  # fmt: skip
    points = np.array([[0, 0, 150], [80, 82, 4000], [10, 10, 450]], dtype=float).T
    assert _conflicting_pairs(tuple(points), tesseroids) == [(0, 0)]


def test_argument_errors_need_no_device():
    """test/test_tesseroid.py:101-131"""
    import harmonica_b200 as hb

    with pytest.raises(ValueError, match="Gravitational field this-field-does-not-exist not recognized"):
        hb.tesseroid_gravity([0, 0, 0], [-10, 10, -10, 10, 100, 200], 1000, "this-field-does-not-exist")
    with pytest.raises(ValueError, match="The bottom radius boundary can't be greater than the top one"):
        hb.tesseroid_gravity([0.0, 0.0, 10.0], [0.0, 10.0, 0.0, 10.0, 20.0, 10.0], 100.0, "potential")


# ------------------------------------------------------------------ tesseroid layer (host logic)
def test_tesseroid_layer_host_logic():
    """test/test_tesseroid_layer.py:85-133, 165-293, 535-560"""
    import harmonica_b200 as hb
    from harmonica_b200._tesseroid_layer import _discard_thin_tesseroids

    R = MEAN_RADIUS
    latitude = np.linspace(-10, 10, 6)
    for west, east in [(0, 480), (0, 360), (-180, 180), (0, 360 - 18 / 2)]:
        longitude = np.linspace(west, east, 21)
        with pytest.raises(ValueError, match="overlapping tesseroids around the globe"):
            hb.tesseroid_layer((longitude, latitude), R * np.ones((6, 21)) + 1e3, R * np.ones((6, 21)))
    for west, east in [(0, 360 - 18), (-180, 180 - 18)]:
        hb.tesseroid_layer((np.linspace(west, east, 21), latitude), R * np.ones((6, 21)) + 1e3, R)
    longitude = np.linspace(-10, 10, 5)
    surface = R * np.ones((6, 5))
    layer = hb.tesseroid_layer((longitude, latitude), surface, surface - 1e3)
    acc = layer.tesseroid_layer
    assert acc.dims == ("latitude", "longitude") and acc.spacing == (4, 5)
    assert acc.boundaries == (longitude[0] - 2.5, longitude[-1] + 2.5, latitude[0] - 2, latitude[-1] + 2)
    assert acc.size == 30 and acc.shape == (6, 5)
    npt.assert_allclose(acc.top, R)
    npt.assert_allclose(acc.bottom, R - 1e3)
    # surface below the reference: top and bottom swap
    acc.update_top_bottom(surface - 2e3, R - 1e3)
    npt.assert_allclose(acc.top, R - 1e3)
    npt.assert_allclose(acc.bottom, R - 2e3)
    with pytest.raises(ValueError, match="Invalid surface array with shape"):
        hb.tesseroid_layer((longitude, latitude), np.ones((7, 5)), R)
    with pytest.raises(ValueError, match="Invalid reference array with shape"):
        hb.tesseroid_layer((longitude, latitude), surface, np.ones((6, 4)))
    with pytest.raises(ValueError, match="Passed longitude coordinates are not evenly spaced"):
        hb.tesseroid_layer((np.array([-10.0, -5, 0, 7, 10]), latitude), surface, R)
    with pytest.raises(ValueError, match="Passed latitude coordinates are not evenly spaced"):
        hb.tesseroid_layer((longitude, np.array([-10.0, -5, 0, 7, 10, 20])), surface, R)
    small = hb.tesseroid_layer((np.linspace(-2, 2, 2), np.linspace(-1, 1, 2)), R * np.ones((2, 2)),
                               (R - 1e3) * np.ones((2, 2)))  # fmt: skip
    expected = [[-4.0, 0.0, -2.0, 0.0, R - 1e3, R], [0.0, 4.0, -2.0, 0.0, R - 1e3, R],
                [-4.0, 0.0, 0.0, 2.0, R - 1e3, R], [0.0, 4.0, 0.0, 2.0, R - 1e3, R]]  # fmt: skip
    npt.assert_allclose(expected, small.tesseroid_layer._to_tesseroids())
    for i in range(2):
        for j in range(2):
            npt.assert_allclose(small.tesseroid_layer.get_tesseroid((i, j)), expected[2 * i + j])
    boundaries = np.array([[-10.0, 10.0, -10.0, 10.0, 0.0, 55.1], [10.0, 30.0, -10.0, 10.0, 0.0, 55.01],
                           [-10.0, 10.0, 10.0, 30.0, 0.0, 35.0], [10.0, 30.0, 10.0, 30.0, 0.0, 84.0]])  # fmt: skip
    thick, rho = _discard_thin_tesseroids(boundaries, np.array([2306, 2122, 2190, 2069]), 55.05)
    npt.assert_allclose(thick, boundaries[[0, 3]])
    npt.assert_allclose(rho, [2306, 2069])
    # NaN masks (:296-383)
    holes = R * np.ones((6, 5)) + 1e3
    holes[3, 3] = holes[2, 1] = np.nan
    density = 2670.0 * np.ones((6, 5))
    density[0, 0] = np.nan
    layer = hb.tesseroid_layer((longitude, latitude), holes, R, properties={"density": density})
    mask = layer.tesseroid_layer._get_nonans_mask()
    assert mask.sum() == 28 and not mask[3, 3] and not mask[2, 1]
    with pytest.warns(UserWarning, match="Found missing values in 'density' property"):
        mask = layer.tesseroid_layer._get_nonans_mask(property_name="density")
    assert mask.sum() == 27 and not mask[0, 0]


# ------------------------------------------------------------------ deferred kernel algorithm
def harness_tesseroid_deferred(coordinates, tesseroids, density, field, radial, defer_cap=16, fast=False,
                               density_upper=None):
    """Host emulation of one thread of tesseroid_deferred_kernel (root records + deferred walks)."""
    H = harness()
    dp = ctypes.POINTER(ctypes.c_double)
    lon, lat, rad = (np.ascontiguousarray(np.atleast_1d(c), dtype=np.float64).ravel() for c in coordinates)
    tesseroids = np.ascontiguousarray(np.atleast_2d(tesseroids), dtype=np.float64)
    density = np.ascontiguousarray(np.atleast_1d(density), dtype=np.float64)
    upper = density if density_upper is None else np.ascontiguousarray(density_upper, dtype=np.float64)
    out = np.zeros(lon.size)
    counts = np.zeros((lon.size, tesseroids.shape[0]), dtype=np.int64)
    flags = ctypes.c_uint(0)
    H.hbt_tesseroid_loop_deferred(
        {"potential": 0, "g_z": 3}[field], ctypes.c_int64(lon.size), lon.ctypes.data_as(dp),
        lat.ctypes.data_as(dp), rad.ctypes.data_as(dp), ctypes.c_int64(tesseroids.shape[0]),
        tesseroids.ctypes.data_as(dp), density.ctypes.data_as(dp), upper.ctypes.data_as(dp),
        int(radial), int(defer_cap),
        int(fast), out.ctypes.data_as(dp), counts.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
        ctypes.byref(flags),
    )  # fmt: skip
    return out, counts, flags.value


@pytest.mark.parametrize("defer_cap", [1, 3, 16])
@pytest.mark.parametrize("field,radial", MODES)
def test_deferred_algorithm_matches_the_oracle(field, radial, defer_cap):
    """root records + deferred walks: the same leaves as the reference for every pair; the sum
    differs only by the order in which a thread adds its pairs (split pairs come later)"""
    g, cases = _cases()
    for name in ("random", "four", "wrapped"):
        coords, tesseroids, density = cases[name]
        tesseroids = np.atleast_2d(np.asarray(tesseroids, dtype=float))
        if (tesseroids[:, 0] > tesseroids[:, 1]).any():
            tesseroids = O.longitude_continuity(tesseroids)
        density = np.atleast_1d(np.asarray(density, dtype=float))
        want, want_counts = O.tesseroid_gravity(coords, tesseroids, density, field, radial, return_counts=True)
        got, counts, flags = harness_tesseroid_deferred(coords, tesseroids, density, field, radial, defer_cap)
        if field == "g_z":
            got *= -1e5
        assert flags == 0
        assert np.array_equal(counts, want_counts)
        want = np.asarray(want).ravel()
        assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))
        if want_counts.max() == 1:  # nothing deferred: the order is the reference's, bit for bit
            assert np.array_equal(got, want)


def test_deferred_algorithm_flags():
    R = MEAN_RADIUS
    _, _, flags = harness_tesseroid_deferred([0, 0, R], [-1.0, 1.0, -1.0, 1.0, R - 1000, R], 2670.0, "g_z", True)
    assert flags & 2  # numba's ZeroDivisionError, found while walking the deferred pair


# ------------------------------------------------------------------ wrapper logic without a device
class _StandInLibrary:
    """The two tesseroid entry points of libharmonica_b200.so answered by the host build of the
    same pair function: lets the CPU suite run the wrapper's own logic (checks, null discard,
    observer ordering, flags -> exceptions, dtype / shape handling)."""

    def hb200_num_devices(self):
        return 1

    @staticmethod
    def _array(pointer, count):
        return np.ctypeslib.as_array(pointer, shape=(count,)).copy() if count else np.zeros(0)

    def hb200_tesseroid_gravity(self, lon, lat, rad, n, tess, rho, n_tess, field, radial, shard, out, flags):
        result = np.ctypeslib.as_array(out, shape=(n,))
        if n_tess == 0:
            result[:] = 0.0
            return 0
        tesseroids = self._array(tess, n_tess * 6).reshape(n_tess, 6)
        got, _, flag_bits = harness_tesseroid_deferred(
            (self._array(lon, n), self._array(lat, n), self._array(rad, n)), tesseroids,
            self._array(rho, n_tess), {0: "potential", 3: "g_z"}[field], radial)  # fmt: skip
        result[:] = got * (-1e5 if field == 3 else 1.0)
        flags._obj.value = flag_bits
        return 0

    def hb200_tesseroid_gravity_variable_density(self, lon, lat, rad, n, tess, rho0, rho1, n_tess, field, shard,
                                                 out, flags):
        result = np.ctypeslib.as_array(out, shape=(n,))
        tesseroids = self._array(tess, n_tess * 6).reshape(n_tess, 6)
        got, _, flag_bits = harness_tesseroid_deferred(
            (self._array(lon, n), self._array(lat, n), self._array(rad, n)), tesseroids,
            self._array(rho0, n_tess), {0: "potential", 3: "g_z"}[field], False,
            density_upper=self._array(rho1, n_tess))  # fmt: skip
        result[:] = got * (-1e5 if field == 3 else 1.0)
        flags._obj.value = flag_bits
        return 0

    def hb200_tesseroid_gravity_density_function(self, lon, lat, rad, n, tess, rho0, rho1, n_tess, field,
                                                 callback, user, out, flags):
        """answered by the oracle, with the density read through the wrapper's C callback (so the
        marshalling of the callback is what this exercises)"""
        from harmonica_b200 import _lib

        function = ctypes.cast(callback, _lib.DENSITY_FN)
        dp = ctypes.POINTER(ctypes.c_double)

        def density(radius):
            radii = np.atleast_1d(np.asarray(radius, dtype=np.float64)).copy()
            values = np.empty_like(radii)
            function(radii.ctypes.data_as(dp), values.ctypes.data_as(dp), radii.size, None)
            return values if np.ndim(radius) else float(values[0])

        tesseroids = self._array(tess, n_tess * 6).reshape(n_tess, 6)
        bottom, top = tesseroids[:, 4], tesseroids[:, 5]
        node = 0.5773502691896257
        for sign, given in ((-1, self._array(rho0, n_tess)), (1, self._array(rho1, n_tess))):
            assert np.array_equal(density(0.5 * (top - bottom) * sign * node + 0.5 * (top + bottom)), given)
        result = np.ctypeslib.as_array(out, shape=(n,))
        result[:] = O.tesseroid_gravity_variable_density(
            (self._array(lon, n), self._array(lat, n), self._array(rad, n)), tesseroids, density,
            {0: "potential", 3: "g_z"}[field], radial_adaptive_discretization=True)
        flags._obj.value = 0
        return 0

    def hb200_tesseroid_inside_scan(self, lon, lat, rad, n, tess, n_tess, flags):
        from harmonica_b200._tesseroid import _conflicting_pairs

        tesseroids = self._array(tess, n_tess * 6).reshape(n_tess, 6)
        pairs = _conflicting_pairs((self._array(lon, n), self._array(lat, n), self._array(rad, n)), tesseroids)
        flags._obj.value = 16 if pairs else 0
        return 0


def test_wrapper_logic_with_a_stand_in_library(monkeypatch):
    import harmonica_b200 as hb
    from harmonica_b200 import _lib

    monkeypatch.setattr(_lib, "ensure_init", lambda: _StandInLibrary())
    rng = np.random.default_rng(5)
    R = MEAN_RADIUS
    w, s = rng.uniform(-60, 50, 12), rng.uniform(-60, 50, 12)
    tesseroids = np.stack([w, w + 8, s, s + 8, np.full(12, R - 2e4), np.full(12, R - 100.0)], axis=1)
    tesseroids[3, 1] = tesseroids[3, 0]  # a null tesseroid
    density = rng.uniform(2000, 3000, 12)
    density[5] = 0.0
    n = 3000  # above the threshold of the locality ordering
    coords = (rng.uniform(-70, 70, n), rng.uniform(-70, 70, n), R + rng.uniform(0, 5e4, n))
    want = O.tesseroid_gravity(coords, tesseroids, density, "g_z")
    ordered = hb.tesseroid_gravity(coords, tesseroids, density, "g_z")
    plain = hb.tesseroid_gravity(coords, tesseroids, density, "g_z", sort_observers=False)
    assert np.array_equal(ordered, plain)  # the order of the observers never changes a value
    assert np.max(np.abs(ordered - want)) <= 1e-13 * np.max(np.abs(want))
    with_bar = hb.tesseroid_gravity(coords, tesseroids, density, "g_z", progressbar=True)
    assert np.array_equal(with_bar, ordered)  # ~20 chunks of computation points, same values
    grid = tuple(c.reshape(50, 60) for c in coords)
    out = hb.tesseroid_gravity(grid, tesseroids, density, "potential", dtype="float32")
    assert out.shape == (50, 60) and out.dtype == np.float32
    with pytest.raises(ValueError, match=re.escape("Found computation point(s) inside tesseroid(s)")):
        hb.tesseroid_gravity([w[0] + 4, s[0] + 4, R - 1e4], tesseroids, density, "g_z")
    with pytest.raises(ZeroDivisionError):
        hb.tesseroid_gravity([0, 0, R], [-1.0, 1.0, -1.0, 1.0, R - 1000, R], 2670.0, "g_z",
                             radial_adaptive_discretization=True)  # fmt: skip
    with pytest.raises(ValueError, match=re.escape("Number of elements in density (2) mismatch")):
        hb.tesseroid_gravity([0, 0, R + 10], tesseroids, [1.0, 2.0], "g_z")


# ------------------------------------------------------------------ fast far field (variant 2)
@pytest.mark.parametrize("field,radial", MODES)
def test_fast_far_field_matches_the_oracle(field, radial):
    """kernel variant 2: arithmetic-only root path (product form of cos(lam_p - lam), reciprocal
    square root, squared thresholds). Same leaves for every pair; values within the parity bar"""
    g, cases = _cases()
    for name in ("random", "four", "wrapped"):
        coords, tesseroids, density = cases[name]
        tesseroids = np.atleast_2d(np.asarray(tesseroids, dtype=float))
        if (tesseroids[:, 0] > tesseroids[:, 1]).any():
            tesseroids = O.longitude_continuity(tesseroids)
        density = np.atleast_1d(np.asarray(density, dtype=float))
        want, want_counts = O.tesseroid_gravity(coords, tesseroids, density, field, radial, return_counts=True)
        got, counts, flags = harness_tesseroid_deferred(coords, tesseroids, density, field, radial, 64, fast=True)
        if field == "g_z":
            got *= -1e5
        assert flags == 0
        assert np.array_equal(counts, want_counts)
        want = np.asarray(want).ravel()
        bar = max(1e-9, 4 * reference_conditioning(coords, tesseroids, density, field, radial, trials=2))
        assert np.max(np.abs(got - want)) <= bar * np.max(np.abs(want))


def test_fast_far_field_accuracy_on_far_and_regional_models():
    """far pairs are well conditioned: the fast path agrees with the oracle to ~1e-13 there; a
    regional model of small tesseroids seen from nearby stays within the conditioning bar"""
    rng = np.random.default_rng(12)
    R = MEAN_RADIUS
    w, s = rng.uniform(-170, 160, 60), rng.uniform(-80, 70, 60)
    tesseroids = np.stack([w, w + rng.uniform(1, 10, 60), s, s + rng.uniform(1, 10, 60),
                           R - rng.uniform(1e3, 1e5, 60), R - rng.uniform(0, 500, 60)], 1)  # fmt: skip
    density = rng.uniform(-1000, 3000, 60)
    coords = (rng.uniform(-180, 180, 80), rng.uniform(-90, 90, 80), R + rng.uniform(2e5, 2e6, 80))
    for field, radial in MODES:
        want, want_counts = O.tesseroid_gravity(coords, tesseroids, density, field, radial, return_counts=True)
        got, counts, flags = harness_tesseroid_deferred(coords, tesseroids, density, field, radial, 64, fast=True)
        if field == "g_z":
            got *= -1e5
        assert flags == 0 and np.array_equal(counts, want_counts)
        assert np.max(np.abs(got - want)) <= 2e-13 * np.max(np.abs(want))
    lon_c, lat_c = np.meshgrid(np.arange(-0.95, 1, 0.1), np.arange(-0.95, 1, 0.1))
    small = np.stack([lon_c.ravel() - 0.05, lon_c.ravel() + 0.05, lat_c.ravel() - 0.05, lat_c.ravel() + 0.05,
                      np.full(lon_c.size, R - 2e3), np.full(lon_c.size, R)], 1)  # fmt: skip
    rho = np.full(lon_c.size, 2670.0)
    near = (rng.uniform(-1, 1, 60), rng.uniform(-1, 1, 60), R + rng.uniform(50, 5e3, 60))
    for field in ("potential", "g_z"):
        want, want_counts = O.tesseroid_gravity(near, small, rho, field, return_counts=True)
        got, counts, flags = harness_tesseroid_deferred(near, small, rho, field, False, 64, fast=True)
        if field == "g_z":
            got *= -1e5
        bar = max(1e-9, 4 * reference_conditioning(near, small, rho, field, False, trials=2))
        assert flags == 0 and np.array_equal(counts, want_counts)
        assert np.max(np.abs(got - want)) <= bar * np.max(np.abs(want)), (field, bar)


# ------------------------------------------------------------------ density given as a function
VD_TOP, VD_BOTTOM = 6371e3, 6371e3 - 3e4


def vd_density_functions():
    """Plain-Python versions of the density functions the fixtures were generated with
    (oracle/make_golden_tesseroid.py: numba-jitted there; same float64 operations)."""
    top, bottom = VD_TOP, VD_BOTTOM
    slope = (2500.0 - 3300.0) / (top - bottom)
    constant_term = 2500.0 - slope * top
    outer, inner, b_factor = 2670.0, 3300.0, 5.0
    a_factor = (inner - outer) / (1 - np.exp(-b_factor))
    exp_constant = inner - a_factor
    thickness = top - bottom
    return {
        "linear": lambda radius: slope * radius + constant_term,
        "exponential": lambda radius: a_factor * np.exp(-b_factor * (radius - bottom) / thickness) + exp_constant,
        "quadratic": lambda radius: 1e-3 * ((radius - 3e3) * (radius - 3e3)) + 1900.0,
    }


@pytest.mark.parametrize("name", ["linear", "exponential"])
def test_variable_density_oracle_is_bit_identical_to_the_reference(name):
    g = golden("tesseroid")
    density = vd_density_functions()[name]
    coords, tesseroids = tuple(g["vd_coords"]), g["vd_tesseroids"]
    disc = O.density_based_discretization(O.longitude_continuity(tesseroids), density)
    assert np.array_equal(disc, g[f"vd_{name}_discretization"])
    for field in ("potential", "g_z"):
        got = O.tesseroid_gravity_variable_density(coords, tesseroids, density, field)
        assert np.array_equal(got, g[f"vd_{name}_{field}"])


def test_density_based_discretization_of_the_product():
    """harmonica_b200/_tesseroid_density.py against the reference's outputs and against the
    closed forms of test/test_tesseroid_variable_density.py:103-320"""
    from harmonica_b200 import _tesseroid_density as D

    g = golden("tesseroid")
    fns = vd_density_functions()
    for name in ("linear", "exponential"):
        disc = D.density_based_discretization(O.longitude_continuity(g["vd_tesseroids"]), fns[name])
        assert np.array_equal(disc, g[f"vd_{name}_discretization"])
    quadratic, bottom, top = fns["quadratic"], 2e3, 5e3
    pieces = np.array(D._density_based_discretization([-3.0, 2.0, -4.0, 5.0, bottom, top], quadratic))
    assert np.array_equal(pieces, g["vd_quadratic_discretization"])
    npt.assert_allclose(D.density_minmax(quadratic, bottom, top), (1900.0, quadratic(top)))
    assert np.array_equal(D.density_minmax(quadratic, bottom, top), g["vd_quadratic_minmax"])
    slope = (quadratic(top) - quadratic(bottom)) / (top - bottom)
    radius_split = 0.5 * slope / 1e-3 + 3e3
    line = lambda radius: slope * (radius - bottom) + quadratic(bottom)  # noqa: E731
    npt.assert_allclose(D.straight_line(3.7e3, quadratic, bottom, top), line(3.7e3))
    npt.assert_allclose(D.maximum_absolute_diff(quadratic, bottom, top),
                        (radius_split, abs(quadratic(radius_split) - line(radius_split))), rtol=1e-6)  # fmt: skip
    # linear and constant densities are never split (:276-320)
    assert len(D._density_based_discretization([-3, 2, -4, 5, 30, 50], lambda radius: 3)) == 1
    assert len(D._density_based_discretization([-3, 2, -4, 5, 30.0, 50.0], lambda radius: 3.1 * radius + 0.4)) == 1
    # the two radial quadrature nodes of every piece
    lower, upper = D.density_at_radial_nodes(pieces, quadratic)
    mid, half = 0.5 * (pieces[:, 5] + pieces[:, 4]), 0.5 * (pieces[:, 5] - pieces[:, 4])
    npt.assert_allclose(lower, [quadratic(r) for r in mid - half / np.sqrt(3)], rtol=1e-15)
    npt.assert_allclose(upper, [quadratic(r) for r in mid + half / np.sqrt(3)], rtol=1e-15)


@pytest.mark.parametrize("fast", [False, True])
@pytest.mark.parametrize("name", ["linear", "exponential"])
def test_variable_density_pair_function_matches_the_reference(name, fast):
    """two densities per tesseroid (host-evaluated at its radial nodes) through the host build of
    the kernel's per-thread algorithm, against the reference's outputs"""
    from harmonica_b200 import _tesseroid_density as D

    g = golden("tesseroid")
    density = vd_density_functions()[name]
    coords = tuple(g["vd_coords"])
    disc = D.density_based_discretization(O.longitude_continuity(g["vd_tesseroids"]), density)
    lower, upper = D.density_at_radial_nodes(disc, density)
    for field in ("potential", "g_z"):
        want = g[f"vd_{name}_{field}"]
        got, _, flags = harness_tesseroid_deferred(coords, disc, lower, field, False, 64, fast=fast,
                                                   density_upper=upper)  # fmt: skip
        if field == "g_z":
            got *= -1e5
        assert flags == 0
        bar = 1e-13 if not fast else 1e-9
        assert np.max(np.abs(got - want)) <= bar * np.max(np.abs(want))
        plain, _, _ = harness_tesseroid(coords, disc, lower, field, False, density_upper=upper)
        if field == "g_z":
            plain *= -1
            plain *= 1e5
        assert np.array_equal(plain, want)  # reference order of additions: bit for bit


def test_variable_density_wrapper_with_a_stand_in_library(monkeypatch):
    """test/test_tesseroid_variable_density.py:322-345 and the wrapper's own rules"""
    import harmonica_b200 as hb
    from harmonica_b200 import _lib

    monkeypatch.setattr(_lib, "ensure_init", lambda: _StandInLibrary())
    g = golden("tesseroid")
    fns = vd_density_functions()
    coords, tesseroids = tuple(g["vd_coords"]), g["vd_tesseroids"]
    for name in ("linear", "exponential"):
        for field in ("potential", "g_z"):
            got = hb.tesseroid_gravity(coords, tesseroids, fns[name], field)
            want = g[f"vd_{name}_{field}"]
            assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))
    # a constant density function gives what the constant density gives
    bottom, top = 5400e3, 6300e3
    tesseroid = [-3, 3, -2, 2, bottom, top]
    lon, lat = np.meshgrid(np.arange(-5, 6, 2.0), np.arange(-5, 6, 2.0))
    grid = (lon, lat, np.full_like(lon, top))
    for field in ("potential", "g_z"):
        npt.assert_allclose(hb.tesseroid_gravity(grid, tesseroid, lambda radius: 2900.0, field),
                            hb.tesseroid_gravity(grid, tesseroid, 2900.0, field))  # fmt: skip
    # ... and with the 3-D discretisation, where the library asks for the density of every leaf
    # through a callback (tesseroid_gravity.py:342-445)
    above = (lon, lat, np.full_like(lon, top + 5e3))
    for field in ("potential", "g_z"):
        npt.assert_allclose(
            hb.tesseroid_gravity(above, tesseroid, lambda radius: 2900.0, field, radial_adaptive_discretization=True),
            hb.tesseroid_gravity(above, tesseroid, 2900.0, field, radial_adaptive_discretization=True), rtol=1e-12)
    g = golden("tesseroid_density_3d")
    got = hb.tesseroid_gravity(tuple(g["coords"]), g["tesseroids"], vd_density_functions()["linear"], "g_z",
                               radial_adaptive_discretization=True)
    npt.assert_allclose(got, g["linear_g_z"], rtol=1e-12)

    def broken(radius):
        raise RuntimeError("density function failed")

    with pytest.raises(RuntimeError, match="density function failed"):
        hb.tesseroid_gravity(above, tesseroid, broken, "g_z", radial_adaptive_discretization=True)


# ------------------------------------------------------------------ own trig in the walks (variant 3)
def harness_trig(op, values):
    H = harness()
    dp = ctypes.POINTER(ctypes.c_double)
    a = np.ascontiguousarray(values, dtype=np.float64)
    out = np.empty_like(a)
    H.hbt_trig(int(op), ctypes.c_int64(a.size), a.ctypes.data_as(dp), out.ctypes.data_as(dp))
    return out


def test_own_trig_sequences_accuracy():
    """hb200_trig.cuh against glibc: sin / cos for |x| < 4 pi and acos on [-1, 1] within 1 ulp;
    acos beyond 1 is NaN like libm's (the split test then reads "no split")"""
    rng = np.random.default_rng(0)

    def ulps(got, want):
        spacing = np.spacing(np.abs(want))
        spacing[spacing == 0] = 5e-324
        return np.abs(got - want) / spacing

    x = np.concatenate([rng.uniform(-4 * np.pi, 4 * np.pi, 400_000), rng.uniform(-1e-3, 1e-3, 50_000),
                        np.arange(-8, 9) * np.pi / 2, np.arange(-8, 9) * np.pi / 2 + 1e-9])  # fmt: skip
    assert ulps(harness_trig(0, x), np.sin(x)).max() <= 1.0
    assert ulps(harness_trig(1, x), np.cos(x)).max() <= 1.0
    c = np.concatenate([rng.uniform(-1, 1, 300_000), 1 - 10.0 ** rng.uniform(-16, -1, 200_000),
                        -1 + 10.0 ** rng.uniform(-16, -1, 100_000),
                        [1.0, -1.0, 0.0, 0.5, -0.5, 1 + 2.3e-16, -1 - 2.3e-16]])  # fmt: skip
    got = harness_trig(2, c)
    with np.errstate(invalid="ignore"):
        want = np.arccos(c)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = np.isfinite(want)
    assert ulps(got[ok], want[ok]).max() <= 1.0


@pytest.mark.parametrize("field,radial", MODES)
def test_own_trig_walks_match_the_oracle(field, radial):
    """kernel variant 3 (host emulation): the walks take sin / cos / acos from hb200_trig.cuh. Same
    leaves as the reference in these models, values within the parity bar"""
    g, cases = _cases()
    for name in ("random", "four", "wrapped"):
        coords, tesseroids, density = cases[name]
        tesseroids = np.atleast_2d(np.asarray(tesseroids, dtype=float))
        if (tesseroids[:, 0] > tesseroids[:, 1]).any():
            tesseroids = O.longitude_continuity(tesseroids)
        density = np.atleast_1d(np.asarray(density, dtype=float))
        want, want_counts = O.tesseroid_gravity(coords, tesseroids, density, field, radial, return_counts=True)
        got, counts, flags = harness_tesseroid_deferred(coords, tesseroids, density, field, radial, 64, fast=2)
        if field == "g_z":
            got *= -1e5
        assert flags == 0
        assert np.array_equal(counts, want_counts)
        want = np.asarray(want).ravel()
        bar = max(1e-9, 4 * reference_conditioning(coords, tesseroids, density, field, radial, trials=2))
        assert np.max(np.abs(got - want)) <= bar * np.max(np.abs(want))


@pytest.mark.parametrize("name", ["linear", "exponential"])
def test_variable_density_3d_oracle_is_bit_identical_to_the_reference(name):
    """density function + radial_adaptive_discretization=True (tesseroid_gravity.py:342-445):
    the oracle against the unmodified reference (tests/golden/tesseroid_density_3d.npz,
    oracle/make_golden_tesseroid_density_3d.py)"""
    g = golden("tesseroid_density_3d")
    density = vd_density_functions()[name]
    for field in ("potential", "g_z"):
        got = O.tesseroid_gravity_variable_density(tuple(g["coords"]), g["tesseroids"], density, field,
                                                   radial_adaptive_discretization=True)
        assert np.array_equal(got, g[f"{name}_{field}"]), field
