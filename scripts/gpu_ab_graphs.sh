#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/g_pytest.log
tail -8 gpurun_out/g_pytest.log
CONFIGS="32:4:1,32:1:1,64:4:1,32:4:2,16:4:1,64:2:1" timeout 300 python scripts/gpu_batch_tune.py > gpurun_out/g_tune.log 2>&1
grep "^B=" gpurun_out/g_tune.log; tail -5 gpurun_out/g_tune.log | grep -v "^B="
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
cat gpurun_out/g_bench.json; tail -n 3 gpurun_out/g_bench.err
