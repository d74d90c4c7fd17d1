"""Run one API case once (for ncu): python profiles/run_case.py acc3|potential [n_prisms] [n_obs]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import harmonica_b200 as hb  # noqa: E402
from _common import config1  # noqa: E402

case = sys.argv[1]
n_p = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
n_o = int(sys.argv[3]) if len(sys.argv) > 3 else 37_888
hb.init([0])
coords, prisms, density = config1(n_p, n_o, seed=1)
fields = {"acc3": ("g_e", "g_n", "g_z"), "potential": "potential", "g_z": "g_z"}[case]
hb.prism_gravity(coords, prisms, density, fields, disable_checks=True)
