#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/sweep_r4p.txt
for cfg in "1 1" "1 2" "0 1" "0 2" "1 8"; do
set -- $cfg
timeout 300 python bench.py --other-configs 0 --steps 10 --no-cpu-baseline --batch-graph $1 --pose-warps $2 > gpurun_out/bench_r4p_bg$1_pw$2.json 2> gpurun_out/bench_r4p_bg$1_pw$2.err
python - <<PY >> gpurun_out/sweep_r4p.txt
import json
d = json.loads(open("gpurun_out/bench_r4p_bg$1_pw$2.json").read().strip().splitlines()[-1])
print("batch_graph $1 pose_warps $2:", round(d["value"], 1), round(d["ms_per_step"], 3), round(d["e2e"]["value"], 1), d["batch_ms"], d["objects_per_frame"], d["single_frame"]["latency_ms"])
PY
done
cat gpurun_out/sweep_r4p.txt
