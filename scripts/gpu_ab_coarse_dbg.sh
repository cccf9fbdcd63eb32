#!/bin/bash
# where is k_match_coarse bound? 0 = shipped kernel, 1 = epilogue returns the accumulator stage unread, 2 = TMEM reads but no scan
mkdir -p gpurun_out
echo "=== shipped"; Q=2000 REPS=20 timeout 300 python scripts/gpu_match_bench.py 2>&1 | grep "tensor\]"
for V in 1 2; do
  echo "=== MC_COARSE_DBG=$V"
  MOPED_LIB=$PWD/moped_b200/lib/libmoped_cuda_dbg$V.so Q=2000 REPS=10 timeout 600 python scripts/gpu_match_bench.py 2>&1 | grep "tensor\]"
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python scripts/real_image_report.py 2>/dev/null | grep "^| bag\|^| ex"
