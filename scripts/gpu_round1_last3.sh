#!/bin/bash
mkdir -p gpurun_out
timeout 18 python -m pytest tests -m gpu -q --timeout=15 -p no:cacheprovider -k "tesseroid" > gpurun_out/pytest_gpu20.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu20.log
tail -3 gpurun_out/pytest_gpu20.log
timeout 12 python scripts/time_tesseroid_order.py > gpurun_out/tess_fast_timing.jsonl 2> gpurun_out/tess_fast_timing.err
cut -c1-200 gpurun_out/tess_fast_timing.jsonl
