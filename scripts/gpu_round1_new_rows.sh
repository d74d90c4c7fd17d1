#!/bin/bash
# One gpurun call: GPU tests (old + new rows), smoke, first timings and one ncu capture of the
# tesseroid kernel. Every step has its own timeout; everything is logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu_new_rows.txt 2>&1
timeout 190 python -m pytest tests -m gpu -q --timeout=60 -p no:cacheprovider > gpurun_out/pytest_gpu16.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu16.log
tail -5 gpurun_out/pytest_gpu16.log
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke16.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke16.log
tail -3 gpurun_out/smoke16.log
timeout 60 python scripts/time_new_rows.py > gpurun_out/new_rows_timing.jsonl 2> gpurun_out/new_rows_timing.err
cat gpurun_out/new_rows_timing.jsonl
timeout 45 ncu --set full --clock-control none --import-source on -k regex:tesseroid_kernel -c 1 -f -o gpurun_out/prof_tess_r1 \
    python -c "
import sys; sys.path[:0]=['.','tests']
import numpy as np, bench, harmonica_b200 as hb
hb.init([0])
wl=bench.make_workload('tess_gz',8192,0,0)
hb.tesseroid_gravity(wl['coords'],wl['tesseroids'],wl['density'],'g_z',disable_checks=True)
" > gpurun_out/ncu_tess.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out/prof_tess_r1.ncu-rep 2>/dev/null
