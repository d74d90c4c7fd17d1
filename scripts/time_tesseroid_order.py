"""Effect of the locality ordering of the observers on the tesseroid kernels (B200)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import bench  # noqa: E402
import harmonica_b200 as hb  # noqa: E402

hb.init([0])
lib = hb._lib.load()
wl = bench.make_workload("tess_gz", 65536, 0, 0)
lon, lat = np.meshgrid(np.linspace(-180, 179, 360), np.linspace(-89.5, 89.5, 180))
grid = (lon.ravel(), lat.ravel(), np.full(lon.size, 6371008.771415059 + 10e3))
for variant in (9, 6, 3):
    lib.hb200_set_tesseroid_variant(variant)
    for name, coords in (("65536 random observers", wl["coords"]), ("360 x 180 grid, row-major", grid)):
        for field in ("g_z", "potential"):
            fn = lambda: hb.tesseroid_gravity(coords, wl["tesseroids"], wl["density"], field,  # noqa: E731
                                              disable_checks=True)
            fn()
            t0 = time.perf_counter()
            fn()
            dt = time.perf_counter() - t0
            print(json.dumps({"row": "tesseroid_gravity " + field, "kernel_variant": variant, "observers": name,
                              "n_obs": int(coords[0].size), "n_tess": int(wl["n_src"]), "seconds": dt,
                              "pairs_per_s": coords[0].size * wl["n_src"] / dt,
                              "api": "numpy host API, e2e (Morton ordering of random observers included)"}),
                  flush=True)
lib.hb200_set_tesseroid_variant(9)
