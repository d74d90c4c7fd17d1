#!/bin/bash
# One gpurun call: GPU tests (old + new rows, both tesseroid kernels), smoke, timings of the new
# rows, one ncu capture of the tesseroid kernel and the tesseroid bench line. Every step has its
# own timeout; everything is logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/gpu_new_rows.txt 2>&1
timeout 120 python -m pytest tests -m gpu -q --timeout=60 -p no:cacheprovider > gpurun_out/pytest_gpu17.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu17.log
tail -6 gpurun_out/pytest_gpu17.log
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke17.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke17.log
tail -4 gpurun_out/smoke17.log
timeout 40 python scripts/time_new_rows.py tesseroids > gpurun_out/new_rows_timing2.jsonl 2> gpurun_out/new_rows_timing2.err
cat gpurun_out/new_rows_timing2.jsonl
timeout 35 ncu --set full --clock-control none --import-source on -k regex:tesseroid_deferred_kernel -c 1 -f -o gpurun_out/prof_tess_r1_v1 \
    python -c "
import sys; sys.path[:0]=['.','tests']
import numpy as np, bench, harmonica_b200 as hb
hb.init([0])
wl=bench.make_workload('tess_gz',8192,0,0)
hb.tesseroid_gravity(wl['coords'],wl['tesseroids'],wl['density'],'g_z',disable_checks=True)
" > gpurun_out/ncu_tess2.log 2>&1
echo "ncu rc=$?"
timeout 70 python bench.py --workload tess_gz --steps 3 --warmup 3 --cpu-seconds 4 > gpurun_out/bench_tess_gz.log 2>&1
echo "bench rc=$?"; tail -1 gpurun_out/bench_tess_gz.log | cut -c1-1500
