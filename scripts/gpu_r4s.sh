#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/sweep_r4s.txt
for sms in 16 8; do
timeout 300 python bench.py --other-configs 0 --steps 10 --no-cpu-baseline --pipeline 1 --stage-sms $sms --lanes 16 > gpurun_out/bench_r4s_p$sms.json 2> gpurun_out/bench_r4s_p$sms.err
python - <<PY >> gpurun_out/sweep_r4s.txt
import json
try:
    d = json.loads(open("gpurun_out/bench_r4s_p$sms.json").read().strip().splitlines()[-1])
    print("pipeline 1 stage_sms $sms:", round(d["value"], 1), round(d["ms_per_step"], 3), round(d["e2e"]["value"], 1), d["objects_per_frame"], d["sm_partition"], d["roofline"]["kernel_ms"])
except Exception as e:
    print("stage_sms $sms failed", e, open("gpurun_out/bench_r4s_p$sms.err").read()[-300:])
PY
done
cat gpurun_out/sweep_r4s.txt
