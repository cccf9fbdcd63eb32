#!/bin/bash
# round 2, session 3, second GPU call: A/B diagnostics of the stream kernel, bucketed refit, the cluster-partitioned frame (one process,
# two / three contexts), the default bench line with its sub-results
mkdir -p gpurun_out; rm -f gpurun_out/fit_stream_ab.txt
timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_batch.py -m gpu -x -q -k "fit_stream or fit_thread or ransac_heavy or cluster_partitioned or sharded" > gpurun_out/pytest_r4b.log 2>&1
echo "pytest rc $?"; tail -5 gpurun_out/pytest_r4b.log; cat gpurun_out/fit_stream_ab.txt
timeout 200 python bench.py --workload ransac --steps 5 --no-cpu-baseline > gpurun_out/bench_r4b_ransac.json 2> gpurun_out/bench_r4b_ransac.err
echo "bench ransac rc $?"; cut -c1-200 gpurun_out/bench_r4b_ransac.json
timeout 200 python bench.py --frames 1 --partition cluster --steps 20 > gpurun_out/bench_r4b_cluster_1gpu.json 2> gpurun_out/bench_r4b_cluster_1gpu.err
echo "bench cluster-partition (1 GPU) rc $?"; cut -c1-300 gpurun_out/bench_r4b_cluster_1gpu.json; tail -3 gpurun_out/bench_r4b_cluster_1gpu.err
timeout 600 python bench.py > gpurun_out/bench_r4b_1gpu.json 2> gpurun_out/bench_r4b_1gpu.err
echo "bench default rc $?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r4b_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
for k, v in (d.get("other_configs") or {}).items():
    print(k, "|", {a: v.get(a) for a in ("value", "unit", "ms_per_step", "wall_s", "error")})
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r4b_ransac.csv \
  python bench.py --workload ransac --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_r4b_ransac.log 2>&1
echo "launch list rc $?"
