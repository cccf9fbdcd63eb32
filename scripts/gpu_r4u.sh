#!/bin/bash
# final state of round 2, session 3: whole GPU suite, smoke, default bench line (with sub-results), reference arm
mkdir -p gpurun_out; rm -f gpurun_out/fit_stream_ab.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r4u.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/pytest_gpu_r4u.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r4u_1gpu.json 2> gpurun_out/bench_r4u_1gpu.err
echo "bench default rc $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r4u_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["batch_ms"], d["single_frame"], d["objects_per_frame"], d["roofline"]["frac"], d["gpu_launches"], d["cpu_baseline"]["value"])
for k, v in (d.get("other_configs") or {}).items():
    print(k, "|", {a: v.get(a) for a in ("value", "unit", "ms_per_step", "wall_s", "error")})
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r4u_reference_arm.json 2> gpurun_out/bench_r4u_reference_arm.err
echo "reference arm rc $?"; cut -c1-300 gpurun_out/bench_r4u_reference_arm.json
