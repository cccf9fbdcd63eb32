#!/bin/bash
# compute-sanitizer memcheck of the kernels added in the third session of round 2, on small shapes (bounded: the box time is short)
mkdir -p gpurun_out
timeout 170 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_stages.py tests/test_gpu_filter_depth.py tests/test_gpu_batch.py -m gpu -q -x \
  -k "(fit_stream and 8-96) or cluster_golden or filter_depth_edge or (cluster_partitioned and 2-0) or batch_graph" 2>&1 | tail -12 > gpurun_out/memcheck_r2b.log
tail -6 gpurun_out/memcheck_r2b.log
timeout 110 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_stages.py tests/test_gpu_filter_depth.py tests/test_gpu_batch.py -m gpu -q -x \
  -k "(fit_stream and 8-96) or cluster_golden or filter_depth_edge or (cluster_partitioned and 2-0)" 2>&1 | tail -8 > gpurun_out/racecheck_r2b.log
tail -4 gpurun_out/racecheck_r2b.log
