#!/bin/bash
# round 2, session 3, third GPU call: the whole GPU suite, stream-kernel occupancy A/B (3 vs 4 CTAs per SM), fused small refit buckets
mkdir -p gpurun_out; rm -f gpurun_out/fit_stream_ab.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r4c.log 2>&1
echo "pytest rc $?"; tail -4 gpurun_out/pytest_gpu_r4c.log; cat gpurun_out/fit_stream_ab.txt
for c in 3 4; do
  timeout 200 python bench.py --workload ransac --steps 10 --no-cpu-baseline --fit-stream-ctas $c > gpurun_out/bench_r4c_ransac_ctas$c.json 2> gpurun_out/bench_r4c_ransac_ctas$c.err
  echo "bench ransac ctas=$c rc $?"; cut -c1-200 gpurun_out/bench_r4c_ransac_ctas$c.json
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r4c_ransac.csv \
  python bench.py --workload ransac --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_r4c_ransac.log 2>&1
echo "launch list rc $?"
