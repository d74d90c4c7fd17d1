#!/bin/bash
# One build -> measure step on a B200 (one gpurun call). Usage:
#   gpurun --timeout 600 -- 'bash scripts/gpu_step.sh TAG [tests] [bench] [ncu_gz] [ncu_tensor] [ncu_mag] [ncu_pot] [sweep]'
# Everything lands under gpurun_out/TAG_*; summaries are made locally from the .ncu-rep files.
TAG=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader > gpurun_out/${TAG}_gpu.txt 2>&1
export_pages() {  # base (without .ncu-rep): CSV pages stay, the 20-40 MB report does not travel
    ncu -i $1.ncu-rep --page raw --csv > $1.raw.csv 2>/dev/null
    ncu -i $1.ncu-rep --page source --csv 2>/dev/null | gzip > $1.source.csv.gz
    ncu -i $1.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip > $1.srcsass.csv.gz
    [ -n "$KEEP_REP" ] || rm -f $1.ncu-rep
}
ncu_case() {  # name kernel-regex python-snippet
    timeout 150 ncu --set full --clock-control none --import-source on -k "regex:$2" -c 1 -f \
        -o gpurun_out/${TAG}_prof_$1 python -c "
import sys; sys.path[:0]=['.','tests']
import numpy as np, bench, harmonica_b200 as hb
hb.init([0])
$3
" > gpurun_out/${TAG}_ncu_$1.log 2>&1
    echo "ncu $1 rc=$?"
    export_pages gpurun_out/${TAG}_prof_$1
}
for what in "$@"; do
case $what in
tests)
    timeout 240 python -m pytest tests -m gpu -q --timeout=90 -p no:cacheprovider -x > gpurun_out/${TAG}_pytest_gpu.log 2>&1
    echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_gpu.log; tail -4 gpurun_out/${TAG}_pytest_gpu.log ;;
smoke)
    timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" ;;
bench)
    timeout 300 python bench.py --cpu-seconds 4 > gpurun_out/${TAG}_bench_default.log 2>&1
    echo "bench rc=$?"; tail -1 gpurun_out/${TAG}_bench_default.log | cut -c1-300 ;;
benchq)  # quick: main workload only
    timeout 120 python bench.py --no-cpu --no-also --no-north-star --steps 3 > gpurun_out/${TAG}_bench_quick.log 2>&1
    echo "benchq rc=$?"; tail -1 gpurun_out/${TAG}_bench_quick.log | cut -c1-200 ;;
tesstests)
    timeout 200 python -m pytest tests/test_gpu_tesseroid.py -m gpu -q --timeout=90 -p no:cacheprovider -x > gpurun_out/${TAG}_pytest_tess.log 2>&1
    echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_tess.log; tail -4 gpurun_out/${TAG}_pytest_tess.log ;;
tilecmp)
    for m in ${TILE_MODES:-0 1}; do
        timeout 120 python bench.py --no-cpu --no-also --no-north-star --steps 3 --tile-mode $m > gpurun_out/${TAG}_bench_tile$m.log 2>&1
        echo "tile mode $m:"; tail -1 gpurun_out/${TAG}_bench_tile$m.log | cut -c1-200
        HB200_TILE_MODE=$m SWEEP_ONLY="prism_gravity g_z" timeout 60 python profiles/field_sweep.py > gpurun_out/${TAG}_sweep_tile$m.jsonl 2>&1
        HB200_TILE_MODE=$m SWEEP_ONLY="fused" timeout 60 python profiles/field_sweep.py >> gpurun_out/${TAG}_sweep_tile$m.jsonl 2>&1
        cut -c1-160 gpurun_out/${TAG}_sweep_tile$m.jsonl
    done ;;
benchref)
    timeout 200 python bench.py --impl reference --steps 3 --warmup 1 --cpu-seconds 4 > gpurun_out/${TAG}_bench_reference.log 2>&1
    echo "benchref rc=$?"; tail -1 gpurun_out/${TAG}_bench_reference.log | cut -c1-200 ;;
launches)
    timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
        --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-also --no-north-star \
        > gpurun_out/${TAG}_launches_bench.log 2>&1; echo "launches rc=$?" ;;
launches_tess)
    timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
        --log-file gpurun_out/${TAG}_launches_tess.csv python bench.py --workload tess_gz --steps 2 --warmup 1 --no-cpu \
        > gpurun_out/${TAG}_launches_tess_bench.log 2>&1; echo "launches_tess rc=$?" ;;
ncu_gz)
    ncu_case gz 'prism_kernel' "
wl=bench.make_workload('layer_gz',37888)
s=wl['sources']
hb.prism_layer_gravity(wl['coords'],s['easting'],s['northing'],s['bottom'],s['top'],s['density'],'g_z')" ;;
ncu_gz_full)  # DRAM traffic of one full-size launch
    timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum \
        --clock-control none -k 'regex:prism_kernel' -c 1 --csv --log-file gpurun_out/${TAG}_traffic_gz_full.csv python -c "
import sys; sys.path[:0]=['.','tests']
import bench, harmonica_b200 as hb
hb.init([0])
wl=bench.make_workload('layer_gz')
s=wl['sources']
hb.prism_layer_gravity(wl['coords'],s['easting'],s['northing'],s['bottom'],s['top'],s['density'],'g_z')" > gpurun_out/${TAG}_traffic.log 2>&1
    echo "traffic rc=$?" ;;
ncu_tensor)
    ncu_case tensor 'prism_kernel' "
wl=bench.make_workload('tensor',37888,100000)
s=wl['sources']
hb.prism_gravity(wl['coords'],s['prisms'],s['density'],wl['fields'],disable_checks=True)" ;;
ncu_mag)
    ncu_case mag 'prism_kernel' "
wl=bench.make_workload('mag_b',37888,100000)
s=wl['sources']
hb.prism_magnetic(wl['coords'],s['prisms'],s['magnetization'],'b',disable_checks=True)" ;;
ncu_pot)
    ncu_case pot 'prism_kernel' "
from _common import config1
c,p,d=config1(20000,37888,seed=1)
hb.prism_gravity(c,p,d,'potential',disable_checks=True)" ;;
ncu_acc3)
    ncu_case acc3 'prism_kernel' "
from _common import config1
c,p,d=config1(20000,37888,seed=1)
hb.prism_gravity(c,p,d,('g_e','g_n','g_z'),disable_checks=True)" ;;
ncu_eqs)
    ncu_case eqs 'point_kernel_cart' "
wl=bench.make_workload('eqs',151552,1000000)
s=wl['sources']
hb.eqs_predict(wl['coords'],s['points'],s['coefs'])" ;;
ncu_tess)
    timeout 200 ncu --set full --clock-control none --import-source on -k "regex:tesseroid_(root|walk|deferred|coop)" -c 2 -f \
        -o gpurun_out/${TAG}_prof_tess python -c "
import sys; sys.path[:0]=['.','tests']
import numpy as np, bench, harmonica_b200 as hb
hb.init([0])
hb._lib.load().hb200_set_tesseroid_variant(int('${TESS_VARIANT:-9}'))
wl=bench.make_workload('tess_gz',65536)
hb.tesseroid_gravity(wl['coords'],wl['tesseroids'],wl['density'],'g_z',disable_checks=True)
" > gpurun_out/${TAG}_ncu_tess.log 2>&1
    echo "ncu tess rc=$?"
    export_pages gpurun_out/${TAG}_prof_tess ;;
benchtess)
    timeout 100 python bench.py --workload tess_gz --steps 3 --warmup 3 --cpu-seconds 4 > gpurun_out/${TAG}_bench_tess_gz.log 2>&1
    echo "benchtess rc=$?"; tail -1 gpurun_out/${TAG}_bench_tess_gz.log | cut -c1-200 ;;
tessorder)
    timeout 60 python scripts/time_tesseroid_order.py > gpurun_out/${TAG}_tess_order_timing.jsonl 2>&1; echo "tessorder rc=$?" ;;
sweep)
    SWEEP_ONLY=${SWEEP_ONLY:-} timeout 150 python profiles/field_sweep.py > gpurun_out/${TAG}_field_sweep.jsonl 2> gpurun_out/${TAG}_field_sweep.err; echo "sweep rc=$?" ;;
*) echo "unknown step $what" ;;
esac
done
echo done
