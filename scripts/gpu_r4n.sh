#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_filter_depth.py tests/test_gpu_stages.py -m gpu -q -k "filter" > gpurun_out/pytest_gpu_r4n.log 2>&1
echo "pytest rc $?"; tail -12 gpurun_out/pytest_gpu_r4n.log
