#!/bin/bash
# round 2, session 3: state to keep — suite, ransac bench + launch list, default bench line with sub-results, stage-kernel ncu
mkdir -p gpurun_out; rm -f gpurun_out/fit_stream_ab.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r4g.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/pytest_gpu_r4g.log; cat gpurun_out/fit_stream_ab.txt
timeout 200 python bench.py --workload ransac --steps 10 > gpurun_out/bench_r4g_ransac.json 2> gpurun_out/bench_r4g_ransac.err
echo "bench ransac rc $?"; cut -c1-200 gpurun_out/bench_r4g_ransac.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r4g_ransac.csv \
  python bench.py --workload ransac --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_r4g_ransac.log 2>&1
echo "launch list rc $?"
timeout 900 python bench.py > gpurun_out/bench_r4g_1gpu.json 2> gpurun_out/bench_r4g_1gpu.err
echo "bench default rc $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r4g_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["batch_ms"], d["single_frame"], d["objects_per_frame"])
for k, v in (d.get("other_configs") or {}).items():
    print(k, "|", {a: v.get(a) for a in ("value", "unit", "ms_per_step", "wall_s", "error")})
PY
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_ransac_first|k_ransac_refit|k_meanshift" -s 40 -c 6 -f -o gpurun_out/prof_stages_r4g \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --other-configs 0 > gpurun_out/prof_stages_r4g.log 2>&1
echo "ncu stages rc $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r4g.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --other-configs 0 > gpurun_out/launches_r4g.log 2>&1
echo "launch list (frames) rc $?"
