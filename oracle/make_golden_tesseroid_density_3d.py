"""
oracle/make_golden_tesseroid_density_3d.py -- TEST INFRASTRUCTURE. Build container only.

tests/golden/tesseroid_density_3d.npz: the reference's UNMODIFIED ``tesseroid_gravity`` (numba)
with a density FUNCTION and ``radial_adaptive_discretization=True``
(_forward/tesseroid_gravity.py:342-445: the function is called at the radial nodes of every leaf
of the 3-D discretisation). Same tesseroids and density functions as the ``vd_*`` entries of
tesseroid.npz; computation points 0.5 .. 200 km above them.
"""
import os
import sys

import numpy as np
from numba import jit

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_shim  # noqa: E402
from make_golden_tesseroid import VD_BOTTOM, VD_TOP, vd_density_functions  # noqa: E402


def main():
    ref = ref_shim.load()
    tg = ref.tesseroid.tesseroid_gravity
    fns = vd_density_functions()
    top, bottom = VD_TOP, VD_BOTTOM
    tesseroids = np.array([[-10, 0, -10, 0, bottom, top], [0, 10, -5, 5, bottom, top - 1e3],
                           [20, 28, 10, 18, bottom + 5e3, top], [350, 5, 20, 30, bottom, top]], dtype=float)  # fmt: skip
    rng = np.random.default_rng(35)
    coords = [rng.uniform(-15, 30, 48), rng.uniform(-15, 32, 48), top + rng.uniform(5e2, 2e5, 48)]
    coords[2][:12] = top + rng.uniform(5e2, 5e3, 12)  # close: the radial direction splits too
    out = {"tesseroids": tesseroids, "coords": np.stack(coords)}
    for name in ("linear", "exponential"):
        for field in ("potential", "g_z"):
            out[f"{name}_{field}"] = np.asarray(
                tg(coords, tesseroids, fns[name], field, parallel=False, radial_adaptive_discretization=True))
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                        "tesseroid_density_3d.npz")
    np.savez(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
