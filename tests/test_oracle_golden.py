"""
The C oracle against tests/golden/*.npz, i.e. against outputs of the
reference's UNMODIFIED wrappers and jitted loops (generated in the build
container by oracle/make_golden.py through oracle/ref_shim.py). CPU only.
Bit-exact: both sides evaluate the same statements with the same libm.
"""

import numpy as np
import numpy.testing as npt
import pytest

import oracle as O
from _common import GRAVITY_FIELDS, golden


@pytest.mark.parametrize("where", ["above", "any"])
@pytest.mark.parametrize("field", GRAVITY_FIELDS)
def test_prism_gravity_random(field, where):
    g = golden("prism_gravity_random")
    coords = tuple(g[f"{where}_{c}"] for c in "enu")
    got = O.prism_gravity(coords, g["prisms"], g["density"], field)
    npt.assert_array_equal(got, g[f"{where}_{field}"])


@pytest.mark.parametrize("field", GRAVITY_FIELDS)
def test_prism_gravity_singular_suite(field):
    g = golden("prism_singular_suite")
    coords = (g["easting"], g["northing"], g["upward"])
    got = O.prism_gravity(coords, g["prisms"][0], g["density"][0], field)
    npt.assert_array_equal(got, g[f"one_{field}"])
    got = O.prism_gravity(coords, g["prisms"], g["density"], field)
    npt.assert_array_equal(got, g[f"two_{field}"])
    if field in GRAVITY_FIELDS[4:]:
        assert np.isnan(got).any() and O.any_singular(coords, g["prisms"], field)


@pytest.mark.parametrize("field", ["b", "b_e", "b_n", "b_u"])
def test_prism_magnetic(field):
    for name, key in (("prism_magnetic_random", ""), ("prism_singular_suite", "two_")):
        g = golden(name)
        coords = (g["easting"], g["northing"], g["upward"])
        got = np.array(O.prism_magnetic(coords, g["prisms"], tuple(g["mag"]), field))
        npt.assert_array_equal(got, g[key + field])


@pytest.mark.parametrize("field", GRAVITY_FIELDS + ("g_ne", "g_ze", "g_zn"))
def test_point_gravity_cartesian(field):
    g = golden("point_gravity_random")
    coords = (g["easting"], g["northing"], g["upward"])
    got = O.point_gravity(coords, tuple(g["points"]), g["masses"], field)
    npt.assert_array_equal(got, g[field])


@pytest.mark.parametrize("field", ["potential", "g_z"])
def test_point_gravity_spherical(field):
    g = golden("point_gravity_random")
    got = O.point_gravity(tuple(g["sph_obs"]), tuple(g["sph_points"]), g["masses"], field, "spherical")
    # numpy's SIMD cos/sin inside numba may differ from glibc's scalar ones in the last ulp
    npt.assert_allclose(got, g[f"sph_{field}"], rtol=1e-12)


@pytest.mark.parametrize("thr", [("thr0", None), ("thr10", 10.0)])
@pytest.mark.parametrize("field", GRAVITY_FIELDS)
def test_prism_layer(field, thr):
    g = golden("prism_layer")
    coords = (g["easting"], g["northing"], g["upward"])
    got = O.prism_layer_gravity(coords, g["east_c"], g["north_c"], g["bottom"], g["top"],
                                g["density"], field, thr[1])
    npt.assert_array_equal(got, g[f"{thr[0]}_{field}"])


def test_eqs_predict_and_jacobian():
    g = golden("eqs_predict")
    coords = (g["easting"], g["northing"], g["upward"])
    npt.assert_array_equal(O.eqs_predict(coords, tuple(g["points"]), g["coefs"]), g["predicted"])
    npt.assert_array_equal(O.eqs_jacobian(coords, tuple(g["points"])), g["jacobian"])


@pytest.mark.parametrize("field", ["b", "b_e", "b_n", "b_u"])
def test_dipole_magnetic(field):
    g = golden("dipole_magnetic")
    coords = (g["easting"], g["northing"], g["upward"])
    got = np.array(O.dipole_magnetic(coords, tuple(g["dipoles"]), tuple(g["moments"]), field))
    npt.assert_array_equal(got, g[field])


def test_eqs_predict_spherical():
    g = golden("eqs_predict_spherical")
    got = O.eqs_predict_spherical(tuple(g["obs"]), tuple(g["points"]), g["coefs"])
    # numpy's SIMD cos/sin inside numba may differ from glibc's scalar ones in the last ulp
    npt.assert_allclose(got, g["predicted"], rtol=1e-11)


def test_numba_loops_match_the_port_bit_for_bit():
    """oracle/numba_loops.py (bench.py's Numba CPU arm) == oracle/choclo_port.c, which the tests
    above pin bit-for-bit on the reference's unmodified loops (golden fixtures)."""
    pytest.importorskip("numba")
    import numba_loops as NL
    from _common import config1, layer_config2

    coords, prisms, density = config1(300, 64, seed=17)
    for field, scale in (("g_z", -1e5), ("potential", 1.0), ("g_ee", 1e9), ("g_ez", -1e9)):
        got = NL.prism_gravity(coords, prisms, density, field) * scale
        np.testing.assert_array_equal(got, O.prism_gravity(coords, prisms, density, field))
    lc, ec, nc, bottom, top, rho = layer_config2(n=24)
    sub = tuple(c[:40] for c in lc)
    got = NL.prism_layer_gravity(sub, ec, nc, bottom, top, rho, "g_z") * -1e5
    np.testing.assert_array_equal(got, O.prism_layer_gravity(sub, ec, nc, bottom, top, rho, "g_z"))
    mag = tuple(np.random.default_rng(3).normal(size=300) for _ in range(3))
    got = np.array(NL.prism_magnetic_field(coords, prisms, mag)) * 1e9
    np.testing.assert_array_equal(got, np.array(O.prism_magnetic(coords, prisms, mag, "b")))
    pts = (prisms[:, 0], prisms[:, 2], prisms[:, 4])
    np.testing.assert_array_equal(NL.eqs_predict(coords, pts, density), O.eqs_predict(coords, pts, density))
