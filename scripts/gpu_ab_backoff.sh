#!/bin/bash
mkdir -p gpurun_out
for NS in 0 40 120 300; do
  echo "=== backoff $NS ns"
  MOPED_LIB=$PWD/moped_b200/lib/libmoped_cuda_bo$NS.so Q=2000 REPS=20 timeout 300 python scripts/gpu_match_bench.py 2>&1 | grep "tensor\]"
  MOPED_LIB=$PWD/moped_b200/lib/libmoped_cuda_bo$NS.so NF=32 timeout 300 python scripts/gpu_match_scale.py 2>&1 | grep "Q= 64000\|Q=  2000"
done
