#!/bin/bash
# A/B of the staged RANSAC kernels against the one-CTA-per-task kernel, then the GPU tests and the bench lines.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/ab_pytest.log
tail -5 gpurun_out/ab_pytest.log
FUSED=1 CONFIGS="32:2:1" timeout 300 python scripts/gpu_batch_tune.py > gpurun_out/ab_fused.log 2>&1
CONFIGS="32:1:1,32:2:1,32:4:1,64:1:1,32:1:2,32:1:4,48:1:1" timeout 300 python scripts/gpu_batch_tune.py > gpurun_out/ab_staged.log 2>&1
echo "--- fused"; grep "^B=" gpurun_out/ab_fused.log
echo "--- staged"; grep "^B=" gpurun_out/ab_staged.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/ab_bench.json 2> gpurun_out/ab_bench.err
timeout 600 python bench.py --workload ransac --steps 5 --no-cpu-baseline > gpurun_out/ab_ransac.json 2> gpurun_out/ab_ransac.err
cat gpurun_out/ab_bench.json gpurun_out/ab_ransac.json
tail -3 gpurun_out/ab_ransac.err gpurun_out/ab_bench.err
