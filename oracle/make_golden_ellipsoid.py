"""
oracle/make_golden_ellipsoid.py -- TEST INFRASTRUCTURE. Build container only.

A reference-held, choclo-free pin for ``dipole_magnetic``: the reference's UNMODIFIED
``ellipsoid_magnetic`` (src/harmonica/_forward/ellipsoids/magnetic.py:37-113; numpy + scipy,
Clark 1986 / Takahashi 2018) evaluated for SPHERES. Outside a uniformly magnetised sphere the
field is exactly that of a dipole of moment 4/3 pi a^3 M at its centre; with an induced
magnetisation M = chi H0 / (1 + chi / 3) (demagnetisation factor 1/3). The reference itself
compares the two at 5e-4 (test/ellipsoids/test_magnetic.py:560-581, without the
demagnetisation term); here the exact relation is stored.

Also writes tests/golden/ellipsoid_sphere_gravity.npz (g_e, g_n, g_z of a homogeneous sphere in
mGal from the reference's ellipsoid_gravity: a point mass outside).

Writes tests/golden/ellipsoid_sphere_magnetic.npz:
  coordinates (3, n), centre (3,), radius, remanent_mag (3,), b_remanent (3, n) [nT],
  susceptibility, inducing_field (3,) [nT], b_induced (3, n) [nT], mu_0 (scipy's, which the
  ellipsoid code uses: CODATA 2022, 1.35e-10 below 4 pi 1e-7)
"""
import os
import sys
import warnings

import numpy as np
from scipy.constants import mu_0

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_shim  # noqa: E402


def main():
    ref = ref_shim.load_ellipsoids()
    rng = np.random.default_rng(7)
    radius, centre = 50.0, (10.0, -20.0, -100.0)
    n = 200
    direction = rng.normal(size=(3, n))
    direction /= np.linalg.norm(direction, axis=0)
    distance = rng.uniform(60.0, 2000.0, n)
    coordinates = np.array([centre[i] + direction[i] * distance for i in range(3)])
    remanent = np.array([1.3, -0.7, 2.1])
    sphere = ref.ellipsoids.Ellipsoid(radius, radius, radius, center=centre, remanent_mag=remanent)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        b_remanent = np.array(ref.magnetic.ellipsoid_magnetic(tuple(coordinates), sphere, (0.0, 0.0, 0.0)))
    chi, inducing = 0.5, np.array([12000.0, -8000.0, -45000.0])
    sphere = ref.ellipsoids.Ellipsoid(radius, radius, radius, center=centre, susceptibility=chi)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        b_induced = np.array(ref.magnetic.ellipsoid_magnetic(tuple(coordinates), sphere, tuple(inducing)))
    # ellipsoid_gravity (gravity.py:30-135) of a homogeneous sphere: exactly a point mass outside.
    # Its only link to choclo is the constant G (choclo.constants.GRAVITATIONAL_CONST, here the
    # value of oracle/choclo_numba.py), so it pins the SIGNS (g_z downward) and the mGal scaling of
    # point_gravity's accelerations, not G.
    density = 2670.0
    sphere = ref.ellipsoids.Ellipsoid(radius, radius, radius, center=centre, density=density)
    g_sphere = np.array(ref.gravity.ellipsoid_gravity(tuple(coordinates), sphere))
    golden_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    np.savez(os.path.join(golden_dir, "ellipsoid_sphere_gravity.npz"), coordinates=coordinates,
             centre=np.array(centre), radius=radius, density=density, g_sphere=g_sphere)
    out = os.path.join(golden_dir, "ellipsoid_sphere_magnetic.npz")
    np.savez(out, coordinates=coordinates, centre=np.array(centre), radius=radius, remanent_mag=remanent,
             b_remanent=b_remanent, susceptibility=chi, inducing_field=inducing, b_induced=b_induced,
             mu_0=mu_0)
    print("wrote", out)


if __name__ == "__main__":
    main()
