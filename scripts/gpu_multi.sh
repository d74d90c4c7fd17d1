#!/bin/bash
# Multi-GPU validation on one box: gpurun --gpus N -- 'bash scripts/gpu_multi.sh TAG N [tests] [bench] [ref] [src] [inproc]'
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv,noheader > gpurun_out/${TAG}_gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for what in "$@"; do
case $what in
tests)
    timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/${TAG}_pytest_multi.log 2>&1
    echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_multi.log; tail -4 gpurun_out/${TAG}_pytest_multi.log ;;
bench)
    timeout 600 $TR bench.py --gpus $N --steps ${STEPS:-20} --warmup 5 > gpurun_out/${TAG}_bench_${N}gpu.log 2>&1
    echo "bench rc=$?"; grep '^{' gpurun_out/${TAG}_bench_${N}gpu.log | tail -1 | cut -c1-400 ;;
bench1)
    timeout 600 python bench.py --gpus 1 --steps ${STEPS:-20} --warmup 5 --no-cpu --no-also > gpurun_out/${TAG}_bench_1gpu.log 2>&1
    echo "bench1 rc=$?"; grep '^{' gpurun_out/${TAG}_bench_1gpu.log | tail -1 | cut -c1-300 ;;
ref)
    timeout 300 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_${N}gpu.log 2>&1
    echo "ref rc=$?"; grep '^{' gpurun_out/${TAG}_bench_reference_${N}gpu.log | tail -1 | cut -c1-700 ;;
src)
    timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 --workload eqs --shard sources --n-obs 4000000 --no-north-star > gpurun_out/${TAG}_bench_eqs_src_${N}gpu.log 2>&1
    echo "src rc=$?"; grep '^{' gpurun_out/${TAG}_bench_eqs_src_${N}gpu.log | tail -1 | cut -c1-300 ;;
weak)
    timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 --scaling weak --no-north-star > gpurun_out/${TAG}_bench_weak_${N}gpu.log 2>&1
    echo "weak rc=$?"; grep '^{' gpurun_out/${TAG}_bench_weak_${N}gpu.log | tail -1 | cut -c1-300 ;;
inproc)
    timeout 300 python profiles/inprocess_multi_gpu.py > gpurun_out/${TAG}_inprocess_multi.jsonl 2> gpurun_out/${TAG}_inprocess_multi.err
    echo "inproc rc=$?"; tail -3 gpurun_out/${TAG}_inprocess_multi.jsonl | cut -c1-300 ;;
esac
done
echo done
