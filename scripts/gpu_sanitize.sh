#!/bin/bash
# compute-sanitizer passes over the kernels added in round 2 (small cases). Usage:
#   gpurun --timeout 900 -- 'bash scripts/gpu_sanitize.sh TAG'
TAG=$1
mkdir -p gpurun_out
cat > /tmp/san_case.py <<'PY'
import sys
sys.path[:0] = [".", "tests", "oracle"]
import numpy as np
import harmonica_b200 as hb
from _common import layer_config2
hb.init([0])
R = 6371008.771415059
lon_c, lat_c = np.meshgrid(np.arange(-179.0, 180.0, 4.0), np.arange(-88.0, 89.0, 4.0))
tess = np.stack([lon_c.ravel() - 2, lon_c.ravel() + 2, lat_c.ravel() - 2, lat_c.ravel() + 2,
                 np.full(lon_c.size, R - 30e3), np.full(lon_c.size, R - 1e3)], axis=1)
rng = np.random.default_rng(1)
rho = rng.uniform(2500, 3300, lon_c.size)
lon = np.concatenate([rng.uniform(-180, 180, 300), [0.0, 10.0]])
lat = np.concatenate([rng.uniform(-85, 85, 300), [89.9, -89.5]])
obs = (lon, lat, np.full(lon.size, R + 10e3))
for field in ("g_z", "potential"):
    for radial in (False, True):
        out = hb.tesseroid_gravity(obs, tess, rho, field, radial_adaptive_discretization=radial, disable_checks=True)
        print("tess", field, radial, float(np.abs(out).max()))
big = np.array([[-60, 60, -60, 60, R - 1000.0, R]])
print("deep", hb.tesseroid_gravity(([0.1], [0.2], [R + 2000.0]), big, [2670.0], "g_z"))
coords, east_c, north_c, bottom, top, density = layer_config2(n=40, seed=7)
sub = tuple(c[::3] for c in coords)
for field in ("g_z", "g_zz", "potential"):
    print("layer", field, float(np.nanmax(np.abs(hb.prism_layer_gravity(sub, east_c, north_c, bottom, top, density, field)))))
PY
for tool in memcheck racecheck; do
    timeout 400 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_case.py > gpurun_out/${TAG}_sanitize_$tool.log 2>&1
    echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/${TAG}_sanitize_$tool.log | tail -3
done
echo done
