#!/bin/bash
# FIRST gpurun call of the next round: the code written after round 1's GPU budget was spent (depth-aware pose stages,
# pose_exact_order) has only been verified on the CPU (tests/test_depth_pose_host.py). Run its GPU tests alone first, then the whole
# suite, then a first timing and a launch list.   gpurun --timeout 1500 -- 'bash scripts/gpu_next_round_first.sh r2a'
TAG=${1:-r2a}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_depth_pose.py -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu_depth_$TAG.log; tail -5 gpurun_out/pytest_gpu_depth_$TAG.log
timeout 300 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_depth_pose.py -m gpu -q -k "explicit_hypotheses and 0" 2>&1 | tail -15 > gpurun_out/racecheck_depth_$TAG.log; tail -3 gpurun_out/racecheck_depth_$TAG.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_$TAG.log; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python scripts/gpu_depth_pose_bench.py > gpurun_out/depth_pose_bench_$TAG.jsonl 2> gpurun_out/depth_pose_bench_$TAG.err; cat gpurun_out/depth_pose_bench_$TAG.jsonl
timeout 300 python bench.py --workload ransac --hyp 256 --steps 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_ransac256_default.json 2>/dev/null; cut -c1-400 gpurun_out/bench_${TAG}_ransac256_default.json
timeout 300 python bench.py --workload ransac --hyp 256 --steps 5 --no-cpu-baseline --pose-mode exact > gpurun_out/bench_${TAG}_ransac256_exact.json 2>/dev/null; cut -c1-400 gpurun_out/bench_${TAG}_ransac256_exact.json
timeout 200 python scripts/gpu_linkage_bench.py > gpurun_out/linkage_bench_$TAG.jsonl 2>&1; cat gpurun_out/linkage_bench_$TAG.jsonl
DEPTH_BENCH_HYP=32 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}_depth_pose.csv python scripts/gpu_depth_pose_bench.py > /dev/null 2>&1
timeout 400 python bench.py > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err; cut -c1-600 gpurun_out/bench_${TAG}_1gpu.json
# one full ncu capture of the depth pose kernels (explicit hypotheses + RANSAC) and of the cached agglomeration, for profiles/
DEPTH_BENCH_HYP=64 timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_depth_hypotheses|k_depth_ransac' -s 2 -c 2 -f \
    -o gpurun_out/prof_depth_pose_$TAG python scripts/gpu_depth_pose_bench.py > gpurun_out/prof_depth_pose_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_link_agglomerate' -s 6 -c 2 -f \
    -o gpurun_out/prof_linkage_$TAG python scripts/gpu_linkage_bench.py > gpurun_out/prof_linkage_$TAG.log 2>&1
ls -la gpurun_out | tail -15
