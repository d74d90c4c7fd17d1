"""In-process multi-GPU timing of the drop-in API (one process, hb200_init over all devices).
    python profiles/inprocess_multi_gpu.py [n_prisms] [n_obs]
Prints one JSON line per device count (1, 2, 4, 8 as available)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import harmonica_b200 as hb  # noqa: E402
from _common import TENSOR_FIELDS, config1  # noqa: E402

n_prisms = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n_obs = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
coords, prisms, density = config1(n_prisms, n_obs, seed=3, scale=10.0)
n_dev = hb._lib.load().hb200_device_count()
ref = None
for n in (8, 4, 2, 1):
    if n > n_dev:
        continue
    hb.init(list(range(n)))
    hb.prism_gravity(tuple(c[:4096] for c in coords), prisms[:4096], density[:4096], "g_z")  # warm-up
    t0 = time.perf_counter()
    ten = hb.prism_gravity(coords, prisms, density, TENSOR_FIELDS, disable_checks=True, shard="observers")
    dt = time.perf_counter() - t0
    ten = np.stack(ten)
    if ref is None:
        ref = ten
    err = float(np.max(np.abs(ten - ref)) / np.max(np.abs(ref)))
    print(json.dumps({"api": "hb.prism_gravity 6 tensor components, in-process observer sharding",
                      "n_gpus": n, "n_prisms": n_prisms, "n_obs": n_obs, "seconds": dt,
                      "pair_per_s": n_prisms * n_obs / dt, "max_rel_diff_vs_first": err}), flush=True)
