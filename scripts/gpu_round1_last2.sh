#!/bin/bash
mkdir -p gpurun_out
timeout 25 python scripts/time_tesseroid_order.py > gpurun_out/tess_order_timing.jsonl 2> gpurun_out/tess_order_timing.err
cut -c1-260 gpurun_out/tess_order_timing.jsonl
timeout 25 python -m pytest tests -m gpu -q --timeout=30 -p no:cacheprovider -k "tesseroid" > gpurun_out/pytest_gpu19.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu19.log
tail -3 gpurun_out/pytest_gpu19.log
