#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r4r.log 2>&1
echo "pytest rc $?"; tail -4 gpurun_out/pytest_gpu_r4r.log
timeout 300 python bench.py --other-configs 0 --steps 10 --no-cpu-baseline > gpurun_out/bench_r4r.json 2> gpurun_out/bench_r4r.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r4r.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), round(d["ms_per_step"], 3), round(d["e2e"]["value"], 1), d["batch_ms"], d["objects_per_frame"], d["single_frame"], d["gpu_launches"], d["config"]["pose_warps_per_task"], d["config"]["frame_lanes"])
PY
