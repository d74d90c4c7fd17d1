#!/bin/bash
# compute-sanitizer over the density-function path with the 3-D discretisation (hb200_tess_leaves.cuh)
TAG=$1
mkdir -p gpurun_out
cat > /tmp/san_density.py <<'PY'
import os, sys
sys.path[:0] = [".", "tests", "oracle"]
import numpy as np
import harmonica_b200 as hb
from _common import golden
from test_tesseroid_host import vd_density_functions
hb.init([0])
g = golden("tesseroid_density_3d")
fn = vd_density_functions()["exponential"]
for cap in (None, "1500"):
    if cap:
        os.environ["HB200_LEAF_CAP"] = cap
    for field in ("g_z", "potential"):
        out = hb.tesseroid_gravity(tuple(g["coords"]), g["tesseroids"], fn, field, radial_adaptive_discretization=True)
        print(cap, field, float(np.max(np.abs(out - g["exponential_" + field])) / np.max(np.abs(out))))
PY
for tool in memcheck racecheck; do
    timeout 300 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_density.py > gpurun_out/${TAG}_sanitize_density_$tool.log 2>&1
    echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/${TAG}_sanitize_density_$tool.log | tail -2
done
echo done
