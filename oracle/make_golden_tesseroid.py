"""
oracle/make_golden_tesseroid.py -- TEST INFRASTRUCTURE. Run in the build container only:

    python oracle/make_golden_tesseroid.py

Writes tests/golden/tesseroid.npz from the reference's UNMODIFIED ``tesseroid_gravity``
(harmonica/_forward/tesseroid_gravity.py + _tesseroid_utils.py, real numba) loaded from
/root/reference through oracle/ref_shim.py: a seeded random model seen from the surface and from
altitude, the reference's doctest tesseroid, the four-tesseroid case of
test/test_tesseroid.py:71-98 and a west > east (longitude continuity) case, each for both fields
and both discretisation modes, plus the leaves ``_adaptive_discretization`` produces for two pairs.
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
MEAN_RADIUS = 6371008.771415059  # boule.WGS84.mean_radius


def random_model(seed=21, n_tess=40, n_obs=60):
    rng = np.random.default_rng(seed)
    R = MEAN_RADIUS
    w = rng.uniform(-30, 25, n_tess)
    s = rng.uniform(-40, 35, n_tess)
    tesseroids = np.stack(
        [w, w + rng.uniform(0.5, 5, n_tess), s, s + rng.uniform(0.5, 5, n_tess),
         R - rng.uniform(1e3, 5e4, n_tess), R - rng.uniform(0, 900, n_tess)], axis=1)  # fmt: skip
    density = rng.uniform(-500, 3000, n_tess)
    coords = [rng.uniform(-35, 35, n_obs), rng.uniform(-45, 45, n_obs), R + rng.uniform(0, 3e5, n_obs)]
    coords[2][:20] = R  # on the reference sphere: many pairs split deeply
    return coords, tesseroids, density


VD_TOP, VD_BOTTOM = 6371e3, 6371e3 - 3e4
VD_LINEAR = (2500.0, 3300.0)                      # density at the top / at the bottom
VD_EXPONENTIAL = (2670.0, 3300.0, 5.0)            # density outer, inner, b factor
VD_QUADRATIC = (1e-3, 3e3, 1900.0, 2e3, 5e3)      # test/test_tesseroid_variable_density.py:45-62


def vd_density_functions():
    """The density families of the reference's tests (test_tesseroid_variable_density.py:
    linear :443-462, exponential :485-510, quadratic :45-100), numba-jitted like there."""
    from numba import jit

    top, bottom = VD_TOP, VD_BOTTOM
    slope = (VD_LINEAR[0] - VD_LINEAR[1]) / (top - bottom)
    constant_term = VD_LINEAR[0] - slope * top

    @jit(nopython=True)
    def linear(radius):
        return slope * radius + constant_term

    outer, inner, b_factor = VD_EXPONENTIAL
    a_factor = (inner - outer) / (1 - np.exp(-b_factor))
    exp_constant = inner - a_factor
    thickness = top - bottom

    @jit(nopython=True)
    def exponential(radius):
        return a_factor * np.exp(-b_factor * (radius - bottom) / thickness) + exp_constant

    factor, vertex_radius, vertex_density = VD_QUADRATIC[:3]

    @jit(nopython=True)
    def quadratic(radius):
        return factor * (radius - vertex_radius) ** 2 + vertex_density

    return {"linear": linear, "exponential": exponential, "quadratic": quadratic}


def variable_density_cases(ref):
    """Outputs of the reference's unmodified variable-density path (real numba-jitted density
    functions): density-based discretisation and tesseroid_gravity, horizontal discretisation."""
    tg = ref.tesseroid.tesseroid_gravity
    vd = ref.tesseroid_variable_density
    fns = vd_density_functions()
    top, bottom = VD_TOP, VD_BOTTOM
    tesseroids = np.array([[-10, 0, -10, 0, bottom, top], [0, 10, -5, 5, bottom, top - 1e3],
                           [20, 28, 10, 18, bottom + 5e3, top], [350, 5, 20, 30, bottom, top]], dtype=float)  # fmt: skip
    rng = np.random.default_rng(33)
    coords = [rng.uniform(-15, 30, 40), rng.uniform(-15, 32, 40), top + rng.uniform(0, 2e5, 40)]
    coords[2][:10] = top  # on the outer surface
    out = {"vd_tesseroids": tesseroids, "vd_coords": np.stack(coords)}
    for name in ("linear", "exponential"):
        out[f"vd_{name}_discretization"] = vd.density_based_discretization(
            ref.tesseroid_utils._longitude_continuity(tesseroids), fns[name])
        for field in ("potential", "g_z"):
            out[f"vd_{name}_{field}"] = np.asarray(tg(coords, tesseroids, fns[name], field, parallel=False))
    # test_tesseroid_variable_density.py:223-273: one tesseroid, quadratic density
    b, t = VD_QUADRATIC[3:]
    out["vd_quadratic_discretization"] = np.array(
        vd._density_based_discretization([-3.0, 2.0, -4.0, 5.0, b, t], fns["quadratic"]))
    out["vd_quadratic_minmax"] = np.array(vd.density_minmax(fns["quadratic"], b, t))
    out["vd_quadratic_max_abs_diff"] = np.array(vd.maximum_absolute_diff(fns["quadratic"], b, t))
    return out


def main():
    ref = ref_shim.load()
    tg = ref.tesseroid.tesseroid_gravity
    data = {}
    coords, tesseroids, density = random_model()
    data.update(random_coords=np.stack(coords), random_tesseroids=tesseroids, random_density=density)
    R = MEAN_RADIUS
    cases = {
        "random": (coords, tesseroids, density),
        # tesseroid_gravity.py:170-183 (doctest): point on the top surface
        "doctest": ([0, 0, R], [-1.0, 1.0, -1.0, 1.0, R - 1000, R], 2670.0),
        # test/test_tesseroid.py:71-98
        "four": ([[-5.0, 0.0, 1.0], [-5.0, 0.0, 5.0], [R + 100] * 3],
                 [[-10.0, 0, -10.0, 0, R - 1e3, R], [-10.0, 0, 0, 10.0, R - 1e3, R],
                  [0, 10.0, -10.0, 0, R - 1e3, R], [0, 10.0, 0, 10.0, R - 1e3, R]],
                 1000.0 * np.ones(4)),
        # test/test_tesseroid.py:444-462: west > east
        "wrapped": ([0, 0, R + 1e3], [350, 10, -10, 10, R - 1e4, R], 1e3),
    }  # fmt: skip
    for name, (c, t, d) in cases.items():
        for field in ("potential", "g_z"):
            for radial in (False, True):
                key = f"{name}_{field}_{'3d' if radial else '2d'}"
                try:
                    data[key] = np.asarray(tg(c, t, d, field, radial_adaptive_discretization=radial,
                                              parallel=False))
                except ZeroDivisionError:
                    # a point ON a tesseroid corner with 3-D discretisation: every level splits
                    # again until a dimension evaluates to exactly 0 and numba's division raises
                    data[key] = np.array("ZeroDivisionError")
    # leaves of single pairs (test/test_tesseroid.py:585-661 geometry)
    ad = ref.tesseroid_utils._adaptive_discretization
    tess = np.array([-10.0, 10.0, -10.0, 10.0, 1.0, 10.0])
    for tag, point, ratio, radial in (("leaves_2d", [0.0, 0.0, 10.0], 10.0, False),
                                      ("leaves_3d", [0.0, 0.0, 10.5], 3.0, True)):
        stack = np.empty((100, 6))
        small = np.empty((100000, 6))
        n = ad(np.array(point), tess, ratio, stack, small, radial)
        data[tag] = small[:n].copy()
        data[tag + "_setup"] = np.array(point + [ratio])
    data.update(variable_density_cases(ref))
    np.savez(os.path.join(OUT, "tesseroid.npz"), **data)
    print("wrote", os.path.join(OUT, "tesseroid.npz"), os.path.getsize(os.path.join(OUT, "tesseroid.npz")), "bytes")
    for k in sorted(data):
        print(" ", k, data[k].shape)


if __name__ == "__main__":
    main()
