#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/sweep_r4t.txt
timeout 300 python -m pytest tests/test_gpu_stages.py tests/test_gpu_batch.py -m gpu -q -x -k "staged or batch_graph or batch_equals" > gpurun_out/pytest_gpu_r4t.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/pytest_gpu_r4t.log
for cfg in "1 1" "1 4" "0 1"; do
set -- $cfg
timeout 300 python bench.py --other-configs 0 --steps 10 --no-cpu-baseline --merge-levels $1 --pose-warps $2 > gpurun_out/bench_r4t_ml$1_pw$2.json 2> gpurun_out/bench_r4t_ml$1_pw$2.err
python - <<PY >> gpurun_out/sweep_r4t.txt
import json
d = json.loads(open("gpurun_out/bench_r4t_ml$1_pw$2.json").read().strip().splitlines()[-1])
print("merge_levels $1 pose_warps $2:", round(d["value"], 1), round(d["ms_per_step"], 3), round(d["e2e"]["value"], 1), d["batch_ms"], d["objects_per_frame"])
PY
done
cat gpurun_out/sweep_r4t.txt
