#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_exact_pipeline.py -m gpu -q -x > gpurun_out/pytest_gpu_r4o.log 2>&1
echo "pytest rc $?"; tail -12 gpurun_out/pytest_gpu_r4o.log
for bg in 1 0; do
timeout 300 python bench.py --other-configs 0 --steps 10 --no-cpu-baseline --batch-graph $bg > gpurun_out/bench_r4o_bg$bg.json 2> gpurun_out/bench_r4o_bg$bg.err
echo "bench bg=$bg rc $?"; python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r4o_bg$bg.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["batch_ms"], d["objects_per_frame"], d["gpu_launches"], d["config"]["frame_lanes"])
PY
done
