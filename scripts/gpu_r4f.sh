#!/bin/bash
# round 2, session 3: rolled-loop stream kernel (instruction-fetch bound), LU out of line — suite, benches, ncu
mkdir -p gpurun_out; rm -f gpurun_out/fit_stream_ab.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r4f.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/pytest_gpu_r4f.log; cat gpurun_out/fit_stream_ab.txt
timeout 200 python bench.py --workload ransac --steps 10 --no-cpu-baseline > gpurun_out/bench_r4f_ransac.json 2> gpurun_out/bench_r4f_ransac.err
echo "bench ransac rc $?"; cut -c1-200 gpurun_out/bench_r4f_ransac.json
timeout 600 python bench.py --other-configs 0 > gpurun_out/bench_r4f_1gpu.json 2> gpurun_out/bench_r4f_1gpu.err
echo "bench default rc $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r4f_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["batch_ms"], d["single_frame"], d["objects_per_frame"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r4f_ransac.csv \
  python bench.py --workload ransac --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_r4f_ransac.log 2>&1
echo "launch list rc $?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_pose_fit_stream -s 3 -c 1 -f -o gpurun_out/prof_fit_stream_r4f \
  python bench.py --workload ransac --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_fit_stream_r4f.log 2>&1
echo "ncu rc $?"
