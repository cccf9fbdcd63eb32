#!/bin/bash
# Round-end evidence on one B200: full GPU test suite, smoke, bench lines of both arms for the BASELINE workload and for
# feature extraction. Outputs land in gpurun_out/ (copied into profiles/ afterwards).
TAG=${1:-r1h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_$TAG.log; cat gpurun_out/pytest_gpu_$TAG.log | tail -3
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference_arm.json 2> gpurun_out/bench_${TAG}_reference_arm.err
timeout 400 python bench.py > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err
timeout 300 python bench.py --workload sift --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_sift_reference_arm.json 2> gpurun_out/bench_${TAG}_sift_reference_arm.err
timeout 300 python bench.py --workload sift > gpurun_out/bench_${TAG}_sift_1gpu.json 2> gpurun_out/bench_${TAG}_sift_1gpu.err
timeout 200 python bench.py --workload sift --frames 1 --steps 50 --no-cpu-baseline > gpurun_out/bench_${TAG}_sift_single_frame_1gpu.json 2>/dev/null
for f in gpurun_out/bench_${TAG}_*.json; do echo "== $f"; cut -c1-900 $f; done
