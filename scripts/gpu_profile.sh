#!/bin/bash
# One gpurun call: knob sweep, ncu launch list of the default bench command, ncu --set full captures of the
# three dominant kernels. Outputs land in gpurun_out/ (scripts/summarise_profiles.py turns them into profiles/).
TAG=${1:-r1c}
mkdir -p gpurun_out
timeout 600 python scripts/gpu_batch_tune.py > gpurun_out/tune_$TAG.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_$TAG.bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_match_coarse -s 3 -c 1 -f -o gpurun_out/prof_coarse_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_coarse_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_pose_ransac|k_meanshift' -s 600 -c 6 -f -o gpurun_out/prof_stages_$TAG \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_stages_$TAG.log 2>&1
ls -la gpurun_out | tail -12
tail -15 gpurun_out/tune_$TAG.log
