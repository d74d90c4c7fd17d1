"""
GPU parity tests of ``tesseroid_gravity``: public API -> ctypes ->
``hb200_tesseroid_gravity`` / ``hb200_tesseroid_inside_scan`` -> kernels, against the golden
fixtures written by the reference's unmodified ``tesseroid_gravity`` and against the oracle.
Tolerance: max|got - want| <= max(1e-9, 4 x the reference's own conditioning) * max|want|. The CUDA
build contracts FMAs and uses CUDA's sin / cos / acos; where quadrature nodes are close to the
computation point the reference's 1 - cos(psi) distance amplifies such last-place differences
(``reference_conditioning`` measures it with the oracle: up to 2e-9 for g_z 100 m above a
tesseroid, below 1e-12 elsewhere). The host build of the same source is bit-identical to the
oracle (tests/test_tesseroid_host.py).
"""

import re

import numpy as np
import numpy.testing as npt
import pytest

import oracle as O
from _common import TOL, max_rel
from _common import golden
from test_tesseroid_host import (MEAN_RADIUS, MODES, _cases, _key, _shell, _shell_analytical,
                                 reference_conditioning, vd_density_functions)

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[9, 6, 3, 2, 1, 0], ids=["group-walks", "two-kernel", "own-trig", "fast", "deferred", "plain"])
def tess_variant(request, hb):
    """the tesseroid kernels: 9 = as 6 with the walks done by groups of 16 lanes on a shared stack;
    6 = root pass and walks as two kernels; 3 = one kernel, root records
    + deferred walks + arithmetic-only far field + the library's own trig in the walks; 2 = as 3
    with CUDA's trig; 1 = without the fast far field; 0 = first build"""
    lib = hb._lib.load()
    default = lib.hb200_get_tesseroid_variant()
    assert lib.hb200_set_tesseroid_variant(request.param) == 0
    yield request.param
    lib.hb200_set_tesseroid_variant(default)


@pytest.mark.parametrize("field,radial", MODES)
@pytest.mark.parametrize("name", ["random", "doctest", "four", "wrapped"])
def test_golden_tesseroid_gravity(hb, tess_variant, name, field, radial):
    g, cases = _cases()
    coords, tesseroids, density = cases[name]
    want = g[_key(name, field, radial)]
    if want.dtype.kind == "U":  # the reference's jitted loop raises (division by zero)
        with pytest.raises(ZeroDivisionError):
            hb.tesseroid_gravity(coords, tesseroids, density, field, radial_adaptive_discretization=radial)
        return
    got = hb.tesseroid_gravity(coords, tesseroids, density, field, radial_adaptive_discretization=radial)
    assert np.shape(got) == want.shape
    bar = max(TOL, 4 * reference_conditioning(coords, tesseroids, density, field, radial))
    assert bar <= 2e-8
    assert max_rel(got, want) <= bar


@pytest.mark.parametrize("field,radial", MODES)
def test_tesseroid_gravity_vs_oracle(hb, tess_variant, field, radial):
    rng = np.random.default_rng(91)
    R = MEAN_RADIUS
    n_tess, n_obs = 700, 900
    w, s = rng.uniform(-170, 160, n_tess), rng.uniform(-85, 75, n_tess)
    tesseroids = np.stack([w, w + rng.uniform(0.1, 10, n_tess), s, s + rng.uniform(0.1, 10, n_tess),
                           R - rng.uniform(1e3, 1e5, n_tess), R - rng.uniform(0, 500, n_tess)], 1)  # fmt: skip
    density = rng.uniform(-1000, 3000, n_tess)
    density[::50] = 0.0  # null tesseroids are discarded on the host
    coords = (rng.uniform(-180, 180, n_obs), rng.uniform(-90, 90, n_obs),
              R + rng.uniform(0, 10.0 ** rng.uniform(1, 6, n_obs)))  # fmt: skip
    want = O.tesseroid_gravity(coords, tesseroids, density, field, radial)
    got = hb.tesseroid_gravity(coords, tesseroids, density, field, radial_adaptive_discretization=radial)
    bar = max(TOL, 4 * reference_conditioning(coords, tesseroids, density, field, radial, trials=2))
    assert bar <= 2e-8
    assert max_rel(got, want) <= bar
    # few observers, many tesseroids: the source list is split over grid.y and reduced
    few = tuple(c[:7] for c in coords)
    got = hb.tesseroid_gravity(few, tesseroids, density, field, radial_adaptive_discretization=radial)
    assert np.max(np.abs(got - np.asarray(want)[:7])) <= bar * np.max(np.abs(want))
    # source-sharded combination (one device: same answer)
    got = hb.tesseroid_gravity(few, tesseroids, density, field, radial_adaptive_discretization=radial,
                               shard="sources")  # fmt: skip
    assert np.max(np.abs(got - np.asarray(want)[:7])) <= bar * np.max(np.abs(want))


@pytest.mark.parametrize("field", ["potential", "g_z"])
def test_spherical_shell(hb, tess_variant, field):
    """test/test_tesseroid.py:683-770: the closed form of a homogeneous shell, 0.1 %"""
    lon, lat = np.meshgrid(np.arange(0, 351, 10.0), np.arange(-90, 91, 10.0))
    coords = (lon, lat, np.full_like(lon, MEAN_RADIUS))
    for thickness in (10, 1e3, 1e5):
        tesseroids = _shell((12, 6), thickness)
        want = _shell_analytical(MEAN_RADIUS, MEAN_RADIUS - thickness, 1000, MEAN_RADIUS)[field]
        got = hb.tesseroid_gravity(coords, tesseroids, 1000 * np.ones((12, 6)), field)
        assert got.shape == lon.shape
        npt.assert_allclose(got, want, rtol=1e-3)
        radius = MEAN_RADIUS + 1e3
        tesseroids = _shell((6, 6), thickness)
        want = _shell_analytical(MEAN_RADIUS, MEAN_RADIUS - thickness, 1000, radius)[field]
        got = hb.tesseroid_gravity([0, 0, radius], tesseroids, 1000 * np.ones(36), field,
                                   radial_adaptive_discretization=True)  # fmt: skip
        npt.assert_allclose(got, want, rtol=1e-3)


def test_errors_shapes_and_dtypes(hb):
    R = MEAN_RADIUS
    tess = [-10.0, 10.0, -10.0, 10.0, R - 1e3, R]
    # test/test_tesseroid.py:112-131: density size
    with pytest.raises(ValueError, match=re.escape("Number of elements in density (3) mismatch")):
        hb.tesseroid_gravity([0, 0, R + 100], [tess, tess], [1.0, 2.0, 3.0], "potential")
    # :247-314: points inside tesseroids (device scan), phased longitudes included
    msg = re.escape("Found computation point(s) inside tesseroid(s)")
    for point in ([0, 0, R - 500], [360, 0, R - 500]):
        with pytest.raises(ValueError, match=msg):
            hb.tesseroid_gravity(point, tess, 1000.0, "g_z")
    with pytest.raises(ValueError, match=msg) as err:
        hb.tesseroid_gravity(([0, 80, 25], [0, 82, 25], [150, 4000, 450]),
                             [[-10, 10, -10, 10, 100, 200], [20, 30, 20, 30, 400, 500],
                              [-50, -40, -30, -20, 100, 500]], [1.0, 1.0, 1.0], "potential")  # fmt: skip
    assert "'(0.0, 0.0, 150.0)' inside tesseroid '(-10.0, 10.0, -10.0, 10.0, 100.0, 200.0)'" in str(err.value)
    assert "'(25.0, 25.0, 450.0)' inside tesseroid '(20.0, 30.0, 20.0, 30.0, 400.0, 500.0)'" in str(err.value)
    # points on the faces are outside (:251-268)
    on_faces = np.array([[0, 0, 250], [20, 0, 150], [0, 0, 200], [0, 0, 100], [-10, 0, 150], [0, 10, 150]], dtype=float).T
    hb._tesseroid.check_points_outside_tesseroids(tuple(on_faces), np.atleast_2d([-10.0, 10, -10, 10, 100, 200]))
    # :223-245: disable_checks lets an inverted tesseroid through; its potential is the opposite
    valid = hb.tesseroid_gravity([0.0, 0.0, 10.0], [0.0, 10.0, 0.0, 10.0, 10.0, 20.0], 100.0, "potential")
    inverted = hb.tesseroid_gravity([0.0, 0.0, 10.0], [0.0, 10.0, 0.0, 10.0, 20.0, 10.0], 100.0,
                                    "potential", disable_checks=True)  # fmt: skip
    npt.assert_allclose(inverted, -valid)
    # shapes, dtype, scalars, empty model
    lon, lat = np.meshgrid(np.linspace(-5, 5, 4), np.linspace(-3, 3, 3))
    out = hb.tesseroid_gravity((lon, lat, np.full_like(lon, R + 1e3)), tess, 2670.0, "g_z", dtype="float32")
    assert out.shape == (3, 4) and out.dtype == np.float32
    want = O.tesseroid_gravity((lon, lat, np.full_like(lon, R + 1e3)), tess, 2670.0, "g_z")
    npt.assert_allclose(out, want, rtol=1e-6)
    assert np.shape(hb.tesseroid_gravity([0, 0, R + 10], tess, 2670.0, "potential")) == ()
    zero = hb.tesseroid_gravity([0, 0, R + 10], [tess], [0.0], "g_z")  # all tesseroids null
    assert float(zero) == 0.0
    lib = hb._lib.load()
    before = lib.hb200_launch_count()
    hb.tesseroid_gravity([0, 0, R + 10], tess, 2670.0, "potential")
    assert lib.hb200_launch_count() - before >= 3  # inside scan (pack + scan), pack + kernel
    assert lib.hb200_set_tesseroid_variant(99) != 0


@pytest.mark.parametrize("field", ["potential", "g_z"])
def test_tesseroid_layer_gravity(hb, field):
    """test/test_tesseroid_layer.py:384-532: the layer's gravity equals tesseroid_gravity of its
    tesseroids; NaN surfaces / densities and thin tesseroids are skipped"""
    R = MEAN_RADIUS
    latitude, longitude = np.linspace(-10, 10, 6), np.linspace(-10, 10, 5)
    surface = R * np.ones((6, 5)) + 1e3
    density = 2670.0 * np.ones((6, 5))
    lon, lat = np.meshgrid(np.arange(-10, 11, 7.0), np.arange(-10, 11, 7.0))
    grid_coords = (lon, lat, np.full_like(lon, R + 11e3))
    layer = hb.tesseroid_layer((longitude, latitude), surface, R, properties={"density": density})
    tesseroids = layer.tesseroid_layer._to_tesseroids()
    want = O.tesseroid_gravity(grid_coords, tesseroids, density.ravel(), field)
    got = layer.tesseroid_layer.gravity(grid_coords, field=field)
    assert got.shape == lon.shape and max_rel(got, want) <= TOL
    # holes in the surface and in the density (with a warning)
    holes = surface.copy()
    holes[3, 3] = holes[2, 1] = np.nan
    keep = np.ones(30, dtype=bool)
    keep[[3 * 5 + 3, 2 * 5 + 1]] = False
    layer = hb.tesseroid_layer((longitude, latitude), holes, R, properties={"density": density})
    want = O.tesseroid_gravity(grid_coords, tesseroids[keep], density.ravel()[keep], field)
    assert max_rel(layer.tesseroid_layer.gravity(grid_coords, field=field), want) <= TOL
    rho = density.copy()
    rho[3, 3] = rho[2, 1] = np.nan
    layer = hb.tesseroid_layer((longitude, latitude), surface, R, properties={"density": rho})
    with pytest.warns(UserWarning, match="Found missing values in 'density' property"):
        got = layer.tesseroid_layer.gravity(grid_coords, field=field)
    assert max_rel(got, want) <= TOL
    # thin tesseroids are discarded
    thin = surface.copy()
    thin[0, :] = R + 10.0
    layer = hb.tesseroid_layer((longitude, latitude), thin, R, properties={"density": density})
    all_t = layer.tesseroid_layer._to_tesseroids()
    sel = (all_t[:, 5] - all_t[:, 4]) >= 100.0
    want = O.tesseroid_gravity(grid_coords, all_t[sel], density.ravel()[sel], field)
    got = layer.tesseroid_layer.gravity(grid_coords, field=field, thickness_threshold=100.0)
    assert max_rel(got, want) <= TOL


@pytest.mark.parametrize("name", ["linear", "exponential"])
def test_golden_tesseroid_gravity_density_function(hb, tess_variant, name):
    """variable-density tesseroids (test/test_tesseroid_variable_density.py): the outputs of the
    reference's unmodified path with numba-jitted density functions"""
    g = golden("tesseroid")
    density = vd_density_functions()[name]
    coords, tesseroids = tuple(g["vd_coords"]), g["vd_tesseroids"]
    for field in ("potential", "g_z"):
        got = hb.tesseroid_gravity(coords, tesseroids, density, field)
        assert max_rel(got, g[f"vd_{name}_{field}"]) <= TOL


def test_progressbar_gives_identical_results(hb):
    """test/test_tesseroid.py:771-818"""
    tesseroids = [[30.3, 50.5, -72.2, -34.2, 6e4, 6.1e4], [30.3, 50.5, 20.1, 32.3, 6.1e4, 6.2e4],
                  [-10.3, 5.3, 20.1, 32.3, 6.2e4, 6.3e4]]  # fmt: skip
    densities = [2000, 3000, 4000]
    lon, lat = np.meshgrid(np.arange(-15, 56, 10.0), np.arange(-80, 41, 10.0))
    coordinates = (lon, lat, np.full_like(lon, 6.5e4))
    for field in ("potential", "g_z"):
        plain = hb.tesseroid_gravity(coordinates, tesseroids, densities, field)
        with_bar = hb.tesseroid_gravity(coordinates, tesseroids, densities, field, progressbar=True)
        npt.assert_allclose(plain, with_bar)


def test_density_function_rules(hb):
    """test/test_tesseroid_variable_density.py:322-345: a constant function equals the constant"""
    bottom, top = 5400e3, 6300e3
    tesseroid = [-3, 3, -2, 2, bottom, top]
    lon, lat = np.meshgrid(np.arange(-5, 6, 1.0), np.arange(-5, 6, 1.0))
    grid = (lon, lat, np.full_like(lon, top))
    for field in ("potential", "g_z"):
        npt.assert_allclose(hb.tesseroid_gravity(grid, tesseroid, lambda radius: 2900.0, field),
                            hb.tesseroid_gravity(grid, tesseroid, 2900.0, field))  # fmt: skip
    above = (lon, lat, np.full_like(lon, top + 2e3))
    for field in ("potential", "g_z"):  # ... with the 3-D discretisation too (leaf radii through the callback)
        npt.assert_allclose(
            hb.tesseroid_gravity(above, tesseroid, lambda radius: 2900.0, field, radial_adaptive_discretization=True),
            hb.tesseroid_gravity(above, tesseroid, 2900.0, field, radial_adaptive_discretization=True), rtol=1e-12)


@pytest.mark.parametrize("name", ["linear", "exponential"])
def test_density_function_with_the_radial_discretisation(hb, name, monkeypatch):
    """tesseroid_gravity.py:342-445 with radial_adaptive_discretization=True: the density function
    is wanted at the radial nodes of every LEAF. Golden values from the unmodified reference
    (oracle/make_golden_tesseroid_density_3d.py); also with a leaf buffer so small that the
    library has to retry with fewer computation points per batch, and a failing function."""
    g = golden("tesseroid_density_3d")
    density = vd_density_functions()[name]
    coords, tesseroids = tuple(g["coords"]), g["tesseroids"]
    for field in ("potential", "g_z"):
        want = g[f"{name}_{field}"]
        got = hb.tesseroid_gravity(coords, tesseroids, density, field, radial_adaptive_discretization=True)
        bar = max(TOL, 4 * reference_conditioning(coords, tesseroids, np.full(len(tesseroids), 2900.0), field, True))
        assert bar <= 2e-8
        assert max_rel(got, want) <= bar, field
    monkeypatch.setenv("HB200_LEAF_CAP", "2000")
    got = hb.tesseroid_gravity(coords, tesseroids, density, "g_z", radial_adaptive_discretization=True)
    assert max_rel(got, g[f"{name}_g_z"]) <= 2e-8
    monkeypatch.delenv("HB200_LEAF_CAP")

    def broken(radius):
        raise RuntimeError("density function failed")

    with pytest.raises(RuntimeError, match="density function failed"):
        hb.tesseroid_gravity(coords, tesseroids, broken, "g_z", radial_adaptive_discretization=True)


@pytest.mark.parametrize("variant", [9, 6, 3])
def test_polar_observers_fill_the_split_lists(hb, variant):
    """Next to a pole EVERY tesseroid of the nearest latitude rings is near (hundreds of split
    roots for one observer, against ~20 elsewhere): in the two-kernel variant the per-chunk lists
    of split pairs overflow and the walk kernel takes over the rest of the chunk; more than
    32 768 observers are processed in several batches. Against the oracle."""
    lib = hb._lib.load()
    R = MEAN_RADIUS
    lon_c, lat_c = np.meshgrid(np.arange(-179.0, 180.0, 2.0), np.arange(-89.0, 90.0, 2.0))
    tess = np.stack([lon_c.ravel() - 1, lon_c.ravel() + 1, lat_c.ravel() - 1, lat_c.ravel() + 1,
                     np.full(lon_c.size, R - 30e3), np.full(lon_c.size, R - 1e3)], axis=1)
    rng = np.random.default_rng(12)
    density = rng.uniform(2500, 3300, lon_c.size)
    lon = np.array([0.0, 33.0, -120.0, 77.7, 10.0, -45.0, 179.0, 5.0])
    lat = np.array([89.9, 89.0, -89.5, 88.0, -87.0, 0.0, 45.0, 90.0])
    coords = (lon, lat, np.full(lon.size, R + 10e3))
    default = lib.hb200_get_tesseroid_variant()
    try:
        assert lib.hb200_set_tesseroid_variant(variant) == 0
        for field in ("g_z", "potential"):
            want = O.tesseroid_gravity(coords, tess, density, field)
            got = hb.tesseroid_gravity(coords, tess, density, field, disable_checks=True)
            assert max_rel(got, want) <= 2e-8, field
        # several observer batches, polar observers in the last one
        n = 40_000
        many = (np.concatenate([rng.uniform(-180, 180, n), lon]),
                np.concatenate([rng.uniform(-60, 60, n), lat]), np.full(n + lon.size, R + 10e3))
        got = hb.tesseroid_gravity(many, tess, density, "g_z", disable_checks=True)
        want = O.tesseroid_gravity(coords, tess, density, "g_z")
        assert max_rel(got[n:], want) <= 2e-8
        idx = np.arange(0, n, 1600)
        sub = tuple(c[idx] for c in many)
        assert max_rel(got[idx], O.tesseroid_gravity(sub, tess, density, "g_z")) <= 2e-8
    finally:
        lib.hb200_set_tesseroid_variant(default)


@pytest.mark.parametrize("variant", [9, 6])
def test_deep_trees(hb, variant):
    """Continental tesseroids seen from 2 km above: the discretisation goes 14 levels deep, beyond
    what the groups of the cooperative walk keep on their shared stacks; such lists are handed to
    the exact depth-first walk inside the same kernel. Against the oracle."""
    lib = hb._lib.load()
    R = MEAN_RADIUS
    tess = np.array([[-60, 60, -60, 60, R - 1000.0, R], [60, 180, -60, 60, R - 1500.0, R - 100]])
    density = np.array([2670.0, 2900.0])
    coords = (np.array([0.1, 13.0, 59.0, 100.0, -70.0]), np.array([0.2, -33.0, 59.5, 10.0, 61.0]),
              np.array([R + 2000, R + 2500.0, R + 3000.0, R + 2000.0, R + 5000.0]))
    default = lib.hb200_get_tesseroid_variant()
    try:
        assert lib.hb200_set_tesseroid_variant(variant) == 0
        for field in ("g_z", "potential"):
            want = O.tesseroid_gravity(coords, tess, density, field)
            got = hb.tesseroid_gravity(coords, tess, density, field)
            assert max_rel(got, want) <= TOL, field
    finally:
        lib.hb200_set_tesseroid_variant(default)


@pytest.mark.parametrize("variant", [9, 6])
def test_many_tesseroids_long_chunks(hb, variant):
    """more than 32 768 tesseroids: the chunks of the two-kernel variants grow beyond 128 records
    (at most 256 chunks), the lists of polar observers overflow into long remainders"""
    lib = hb._lib.load()
    R = MEAN_RADIUS
    lon_c, lat_c = np.meshgrid(np.arange(-179.5, 180.0, 1.0), np.arange(-89.5, 90.0, 1.0))
    tess = np.stack([lon_c.ravel() - 0.5, lon_c.ravel() + 0.5, lat_c.ravel() - 0.5, lat_c.ravel() + 0.5,
                     np.full(lon_c.size, R - 20e3), np.full(lon_c.size, R - 2e3)], axis=1)
    rng = np.random.default_rng(21)
    density = rng.uniform(2500, 3300, lon_c.size)
    lon = np.concatenate([rng.uniform(-180, 180, 12), [0.0, 77.0, -100.0]])
    lat = np.concatenate([rng.uniform(-80, 80, 12), [89.8, -89.9, 89.2]])
    coords = (lon, lat, np.full(lon.size, R + 5e3))
    default = lib.hb200_get_tesseroid_variant()
    try:
        assert lib.hb200_set_tesseroid_variant(variant) == 0
        for field in ("g_z", "potential"):
            want = O.tesseroid_gravity(coords, tess, density, field)
            got = hb.tesseroid_gravity(coords, tess, density, field, disable_checks=True)
            assert max_rel(got, want) <= 2e-8, field
    finally:
        lib.hb200_set_tesseroid_variant(default)
