#!/bin/bash
# Round 2, second GPU call: full -m gpu suite (exact-order staged RANSAC, exact pipeline, adaptive matcher on the device), then the
# price of exact-order mode in the frame pipeline and the effect of MATCH chunking with high-priority lanes.
TAG=${1:-r2b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu_$TAG.log; tail -12 gpurun_out/pytest_gpu_$TAG.log
for cfg in "default:" "exact:--pose-mode exact" "chunks2:--chunks 2" "chunks4:--chunks 4" "exact_chunks4:--pose-mode exact --chunks 4"; do
  name=${cfg%%:*}; flags=${cfg#*:}
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $flags > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${TAG}_$name.json"))
    print("$name", round(d["value"],1), "fps  e2e", round(d["e2e"]["value"],1), " step ms", round(d["ms_per_step"],2), d["batch_ms"], "single", d["single_frame"]["latency_ms"], d["single_frame"]["stage_ms"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/bench_${TAG}_$name.err").read()[-1500:])
PY
done
timeout 300 python bench.py --workload ransac --steps 5 --no-cpu-baseline --pose-mode exact --depth-team 8 > gpurun_out/bench_${TAG}_ransac_exact8.json 2>/dev/null; cut -c1-200 gpurun_out/bench_${TAG}_ransac_exact8.json
ls gpurun_out | tail -5
