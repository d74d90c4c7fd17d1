#!/bin/bash
# First gpurun call of the next round (one B200, ~6 minutes): everything that could not be measured
# after the round-1 GPU budget ended. Each step has its own timeout; logs under gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 420 -- 'bash scripts/gpu_round2_start.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/r2_gpu.txt 2>&1
# 1. the whole GPU suite (variable-density tesseroids and the progress-bar test have never run on a GPU)
timeout 150 python -m pytest tests -m gpu -q --timeout=60 -p no:cacheprovider > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_gpu.log; tail -5 gpurun_out/r2_pytest_gpu.log
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke.log
# 2. bench lines: default, tesseroids (variant 2 has no bench line yet)
timeout 100 python bench.py > gpurun_out/r2_bench_default.log 2>&1; tail -1 gpurun_out/r2_bench_default.log | cut -c1-400
timeout 60 python bench.py --workload tess_gz --steps 3 --warmup 3 --cpu-seconds 4 > gpurun_out/r2_bench_tess_gz.log 2>&1
tail -1 gpurun_out/r2_bench_tess_gz.log | cut -c1-400
# 3. ncu: tesseroid variant 2 with ordered observers (the wrapper orders them), and the fit kernels
timeout 40 ncu --set full --clock-control none --import-source on -k regex:tesseroid_deferred_kernel -c 1 -f \
    -o gpurun_out/r2_prof_tess_v2 python -c "
import sys; sys.path[:0]=['.','tests']
import bench, harmonica_b200 as hb
hb.init([0])
wl=bench.make_workload('tess_gz',65536,0,0)
hb.tesseroid_gravity(wl['coords'],wl['tesseroids'],wl['density'],'g_z',disable_checks=True)
" > gpurun_out/r2_ncu_tess.log 2>&1
timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_fit.csv \
    python -c "
import sys; sys.path[:0]=['.','tests']
import numpy as np, harmonica_b200 as hb
hb.init([0])
rng=np.random.default_rng(2); n=8192
c=(rng.uniform(0,50e3,n),rng.uniform(0,50e3,n),rng.uniform(0,500,n))
d=hb.eqs_predict(c,(rng.uniform(0,50e3,50),rng.uniform(0,50e3,50),np.full(50,-5e3)),rng.normal(size=50))
hb.EquivalentSources(depth=1500,damping=1e-3).fit(c,d)
" > gpurun_out/r2_ncu_fit.log 2>&1
timeout 30 python scripts/time_new_rows.py > gpurun_out/r2_new_rows_timing.jsonl 2> gpurun_out/r2_new_rows_timing.err
timeout 20 python scripts/time_tesseroid_order.py > gpurun_out/r2_tess_order_timing.jsonl 2>&1
echo done
