#!/bin/bash
# ncu of the stream kernel after the LDL^T change
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_pose_fit_stream -s 3 -c 1 -f -o gpurun_out/prof_fit_stream_r4e \
  python bench.py --workload ransac --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_fit_stream_r4e.log 2>&1
echo "ncu rc $?"
