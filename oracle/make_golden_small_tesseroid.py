"""
oracle/make_golden_small_tesseroid.py -- TEST INFRASTRUCTURE. Build container only.

A reference-held check of the SCALE, SIGN and UNITS of the prism potential (which no reference
test pins with an absolute value): the reference's UNMODIFIED ``tesseroid_gravity`` (real numba
code, no choclo) for a tesseroid of 0.001 x 0.001 degrees x 100 m at the equator, which is a
111 x 111 x 100 m prism up to the curvature (1e-5) and up to the accuracy of the tesseroid
quadrature itself (0.1 %, its design target). Writes tests/golden/small_tesseroid.npz.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_shim  # noqa: E402


def main():
    ref = ref_shim.load()
    radius = 6371008.771415059
    dlam = dphi = 0.001
    thickness = 100.0
    tesseroid = np.array([[-dlam / 2, dlam / 2, -dphi / 2, dphi / 2, radius - thickness, radius]])
    density = np.array([2670.0])
    rng = np.random.default_rng(3)
    n = 30
    lon, lat = rng.uniform(-0.004, 0.004, n), rng.uniform(-0.004, 0.004, n)
    rad = radius + rng.uniform(50, 400, n)
    potential = ref.tesseroid.tesseroid_gravity((lon, lat, rad), tesseroid, density, "potential")
    g_z = ref.tesseroid.tesseroid_gravity((lon, lat, rad), tesseroid, density, "g_z")
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                       "small_tesseroid.npz")
    np.savez(out, longitude=lon, latitude=lat, radius=rad, tesseroid=tesseroid, density=density,
             mean_radius=radius, potential=potential, g_z=g_z)
    print("wrote", out)


if __name__ == "__main__":
    main()
