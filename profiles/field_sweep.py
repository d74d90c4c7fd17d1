"""Throughput of every entry point / field through the public numpy API (end to end).
    python profiles/field_sweep.py > gpurun_out/field_sweep.jsonl
One JSON line per case: pairs, best-of-3 seconds, pair/s."""
import json
import os
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import harmonica_b200 as hb  # noqa: E402
from _common import GRAVITY_FIELDS, TENSOR_FIELDS, config1, layer_config2  # noqa: E402

hb.init([0])
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 2
hb._lib.load().hb200_set_variant(variant)
if os.environ.get("HB200_TILE_MODE"):
    hb._lib.load().hb200_set_tile_mode(int(os.environ["HB200_TILE_MODE"]))
rng = np.random.default_rng(0)


ONLY = os.environ.get("SWEEP_ONLY", "")


def timeit(name, pairs, fn):
    if ONLY and ONLY not in name:
        return
    fn()
    best = min(_t(fn) for _ in range(3))
    print(json.dumps({"case": name, "variant": variant, "pairs": pairs, "seconds": best,
                      "pair_per_s": pairs / best}), flush=True)


def _t(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


n_p, n_o = 20_000, 262_144
coords, prisms, density = config1(n_p, n_o, seed=1)
pairs = float(n_p) * n_o
for f in GRAVITY_FIELDS:
    timeit(f"prism_gravity {f}", pairs, lambda f=f: hb.prism_gravity(coords, prisms, density, f, disable_checks=True))
timeit("prism_gravity (g_e, g_n, g_z) fused", pairs,
       lambda: hb.prism_gravity(coords, prisms, density, ("g_e", "g_n", "g_z"), disable_checks=True))
timeit("prism_gravity 6 tensor components fused", pairs,
       lambda: hb.prism_gravity(coords, prisms, density, TENSOR_FIELDS, disable_checks=True))
mag = tuple(rng.normal(size=n_p) for _ in range(3))
for f in ("b", "b_e", "b_n", "b_u"):
    timeit(f"prism_magnetic {f}", pairs, lambda f=f: hb.prism_magnetic(coords, prisms, mag, f, disable_checks=True))
# observers ON the top faces of a flat-topped model: every pair takes the rule-exact direct path
flat = prisms.copy()
flat[:, 5] = 0.0
flat[:, 4] = -np.abs(prisms[:, 4])
on_top = (coords[0][:32768], coords[1][:32768], np.zeros(32768))
timeit("prism_gravity g_z, all pairs on the direct path (observers in the top-face plane)",
       float(n_p) * 32768, lambda: hb.prism_gravity(on_top, flat, density, "g_z", disable_checks=True))
# tensor component on the same geometry: every pair takes the rule-exact path (face rule)
timeit("prism_gravity g_zz, all pairs on the exact path (observers in the top-face plane)",
       float(n_p) * 32768, lambda: hb.prism_gravity(on_top, flat, density, "g_zz", disable_checks=True))
timeit("prism_magnetic b, all pairs on the exact path (observers in the top-face plane)",
       float(n_p) * 32768, lambda: hb.prism_magnetic(on_top, flat, mag, "b", disable_checks=True))
lc, east_c, north_c, bottom, top, rho = layer_config2(n=300, seed=2)
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    timeit("prism_layer.gravity g_z 300x300 layer x 90k observers", 90_000.0 * 90_000,
           lambda: hb.prism_layer_gravity(lc, east_c, north_c, bottom, top, rho, "g_z"))
n_s = 400_000
pts = (rng.uniform(-5e4, 5e4, n_s), rng.uniform(-5e4, 5e4, n_s), rng.uniform(-5e3, -1e3, n_s))
w = rng.uniform(1e6, 1e9, n_s)
pairs = float(n_s) * n_o
for f in GRAVITY_FIELDS:
    timeit(f"point_gravity {f}", pairs, lambda f=f: hb.point_gravity(coords, pts, w, f))
timeit("eqs_predict", pairs, lambda: hb.eqs_predict(coords, pts, w))
mom = tuple(rng.normal(size=n_s) for _ in range(3))
for f in ("b", "b_u"):
    timeit(f"dipole_magnetic {f}", pairs, lambda f=f: hb.dipole_magnetic(coords, pts, mom, f))
sph_p = (rng.uniform(-40, 40, 50_000), rng.uniform(-60, 60, 50_000), rng.uniform(6.2e6, 6.3e6, 50_000))
sph_o = (rng.uniform(-45, 45, n_o), rng.uniform(-65, 65, n_o), rng.uniform(6.4e6, 6.5e6, n_o))
for f in ("potential", "g_z"):
    timeit(f"point_gravity spherical {f}", 50_000.0 * n_o,
           lambda f=f: hb.point_gravity(sph_o, sph_p, w[:50_000], f, coordinate_system="spherical"))
timeit("eqs_predict spherical", 50_000.0 * n_o,
       lambda: hb.eqs_predict(sph_o, sph_p, w[:50_000], coordinate_system="spherical"))
jo = tuple(c[:16384] for c in coords)
jp = tuple(p[:8192] for p in pts)
timeit("eqs_jacobian 16384 x 8192 (1 GiB matrix to host)", 16384.0 * 8192, lambda: hb.eqs_jacobian(jo, jp))
