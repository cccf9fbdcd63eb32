#!/bin/bash
# Round 2, first GPU call: the whole -m gpu suite (now with the promoted depth-pose / exact-order cases, the 1M-row parity test and
# the any-DescriptorSize test), the default bench line, the per-hypothesis LM parity table, the price of the exact-order LM on
# configs[3], first timings of the row-f4 kernels, a racecheck pass over the small-shape suites, a launch list of the bench step.
#   gpurun --timeout 1700 -- 'bash scripts/gpu_r2a.sh r2a'
TAG=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_gpu_$TAG.log; tail -6 gpurun_out/pytest_gpu_$TAG.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err; cut -c1-900 gpurun_out/bench_${TAG}_1gpu.json
timeout 300 python scripts/lm_parity_table.py --out gpurun_out/lm_parity_$TAG 2>&1 | tail -8
timeout 300 python bench.py --workload ransac --steps 5 --no-cpu-baseline > gpurun_out/bench_${TAG}_ransac_default.json 2>/dev/null; cut -c1-300 gpurun_out/bench_${TAG}_ransac_default.json
timeout 300 python bench.py --workload ransac --steps 3 --no-cpu-baseline --pose-mode exact > gpurun_out/bench_${TAG}_ransac_exact32.json 2>/dev/null; cut -c1-300 gpurun_out/bench_${TAG}_ransac_exact32.json
timeout 300 python bench.py --workload ransac --steps 3 --no-cpu-baseline --pose-mode exact --depth-team 8 > gpurun_out/bench_${TAG}_ransac_exact8.json 2>/dev/null; cut -c1-300 gpurun_out/bench_${TAG}_ransac_exact8.json
timeout 300 python scripts/gpu_depth_pose_bench.py > gpurun_out/depth_pose_bench_$TAG.jsonl 2> gpurun_out/depth_pose_bench_$TAG.err; cat gpurun_out/depth_pose_bench_$TAG.jsonl
timeout 200 python scripts/gpu_linkage_bench.py > gpurun_out/linkage_bench_$TAG.jsonl 2>&1; cat gpurun_out/linkage_bench_$TAG.jsonl
# racecheck / memcheck on the small-shape suites: the warp-team __syncwarp(mask) code, mean-shift, the staged RANSAC kernels
timeout 500 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_depth_pose.py tests/test_gpu_stages.py -m gpu -q -x \
    -k "explicit_hypotheses or team_width or cached_agglomeration or cluster_golden or pose_ransac_matches or filter_golden or process_frame_recovers" 2>&1 | tail -25 > gpurun_out/racecheck_$TAG.log; tail -4 gpurun_out/racecheck_$TAG.log
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_match.py tests/test_gpu_stages.py -m gpu -q -x \
    -k "ragged_sizes or other_descriptor or cluster_random or pose_too_few or filter_vs_oracle or process_frame_recovers" 2>&1 | tail -25 > gpurun_out/memcheck_$TAG.log; tail -4 gpurun_out/memcheck_$TAG.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -20
