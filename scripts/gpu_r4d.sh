#!/bin/bash
# round 2, session 3, fourth GPU call: LDL^T fast path of the default LM — the whole GPU suite, the LM parity table, benches
mkdir -p gpurun_out; rm -f gpurun_out/fit_stream_ab.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r4d.log 2>&1
echo "pytest rc $?"; tail -8 gpurun_out/pytest_gpu_r4d.log; cat gpurun_out/fit_stream_ab.txt
timeout 300 python scripts/lm_parity_table.py --out gpurun_out/lm_parity_r4d > gpurun_out/lm_parity_r4d.log 2>&1
echo "parity table rc $?"; cat gpurun_out/lm_parity_r4d.md | tail -8
timeout 200 python bench.py --workload ransac --steps 10 --no-cpu-baseline > gpurun_out/bench_r4d_ransac.json 2> gpurun_out/bench_r4d_ransac.err
echo "bench ransac rc $?"; cut -c1-200 gpurun_out/bench_r4d_ransac.json
timeout 600 python bench.py --other-configs 0 > gpurun_out/bench_r4d_1gpu.json 2> gpurun_out/bench_r4d_1gpu.err
echo "bench default rc $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r4d_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["batch_ms"], d["single_frame"], d["objects_per_frame"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r4d_ransac.csv \
  python bench.py --workload ransac --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_r4d_ransac.log 2>&1
echo "launch list rc $?"
