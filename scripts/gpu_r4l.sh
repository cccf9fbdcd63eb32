#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r4l.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/pytest_gpu_r4l.log
timeout 300 python bench.py --other-configs 0 --steps 10 > gpurun_out/bench_r4l_1gpu.json 2> gpurun_out/bench_r4l_1gpu.err
echo "bench rc $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r4l_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["match_db_splits"], d["match_tiers"])
PY
