#!/bin/bash
# Last short gpurun call of round 1: the tests touched since the previous call (tesseroid kernels
# with the larger deferral list, the gradient-boosting comparison), timings of both tesseroid
# kernels and one ncu capture of the default one.
mkdir -p gpurun_out
timeout 65 python -m pytest tests -m gpu -q --timeout=40 -p no:cacheprovider -k "tesseroid or gradient_boost" > gpurun_out/pytest_gpu18.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu18.log
tail -4 gpurun_out/pytest_gpu18.log
timeout 25 python scripts/time_new_rows.py tesseroids > gpurun_out/new_rows_timing3.jsonl 2> gpurun_out/new_rows_timing3.err
grep -v potential gpurun_out/new_rows_timing3.jsonl | cut -c1-230
timeout 25 ncu --set full --clock-control none --import-source on -k regex:tesseroid_deferred_kernel -c 1 -f -o gpurun_out/prof_tess_r1_v1b \
    python -c "
import sys; sys.path[:0]=['.','tests']
import numpy as np, bench, harmonica_b200 as hb
hb.init([0])
wl=bench.make_workload('tess_gz',8192,0,0)
hb.tesseroid_gravity(wl['coords'],wl['tesseroids'],wl['density'],'g_z',disable_checks=True)
" > gpurun_out/ncu_tess3.log 2>&1
echo "ncu rc=$?"
