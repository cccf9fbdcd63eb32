#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_stages.py tests/test_gpu_filter_depth.py -m gpu -q -x > gpurun_out/pytest_gpu_r4q.log 2>&1
echo "pytest rc $?"; tail -4 gpurun_out/pytest_gpu_r4q.log
timeout 300 python bench.py --other-configs 0 --steps 10 --no-cpu-baseline --pose-warps 1 > gpurun_out/bench_r4q.json 2> gpurun_out/bench_r4q.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r4q.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), round(d["ms_per_step"], 3), round(d["e2e"]["value"], 1), d["batch_ms"], d["objects_per_frame"], d["single_frame"])
PY
