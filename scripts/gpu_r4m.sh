#!/bin/bash
# ncu --set full of the 8-bit coarse kernel on a 125 k-row shard (the per-rank MATCH of an 8-GPU job)
mkdir -p gpurun_out
DBG_OBJECTS=125 DBG_CONFIGS=1:0:1 DBG_SPLITS=0 DBG_ITERS=4 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_match_coarse -s 2 -c 1 -f -o gpurun_out/prof_coarse_125k_r4m \
  python scripts/gpu_coarse_dbg.py ncu_125k > gpurun_out/prof_coarse_125k_r4m.log 2>&1
echo "ncu rc $?"
