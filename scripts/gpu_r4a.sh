#!/bin/bash
# round 2, session 3, first GPU call: the persistent phase-synchronous fit kernel (A/B against the one-launch kernel), its ncu capture
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "fit_stream or fit_thread or ransac_heavy" > gpurun_out/pytest_r4a.log 2>&1
echo "pytest rc $?"; tail -5 gpurun_out/pytest_r4a.log
for fs in 1 0; do
  timeout 200 python bench.py --workload ransac --steps 5 --no-cpu-baseline --fit-stream $fs > gpurun_out/bench_r4a_ransac_fs$fs.json 2> gpurun_out/bench_r4a_ransac_fs$fs.err
  echo "bench fs=$fs rc $?"; cut -c1-260 gpurun_out/bench_r4a_ransac_fs$fs.json
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_pose_fit_stream -s 3 -c 1 -f -o gpurun_out/prof_fit_stream_r4a \
  python bench.py --workload ransac --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_fit_stream_r4a.log 2>&1
echo "ncu rc $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r4a_ransac.csv \
  python bench.py --workload ransac --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_r4a_ransac.log 2>&1
echo "launch list rc $?"
