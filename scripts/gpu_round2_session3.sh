#!/bin/bash
# What the third session of round 2 ran on a B200 (one GPU unless noted), in one place. Outputs land in gpurun_out/ and were copied to
# profiles/*_r2b*. Each block is independent; run the whole script or paste a block into `gpurun -- '...'`.
TAG=${1:-r2b}
mkdir -p gpurun_out

# 1. whole GPU suite, smoke, both arms of the default bench line (the default line carries the other BASELINE configs as sub-results)
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_1gpu.json 2> gpurun_out/bench_${TAG}_1gpu.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_reference_arm.json 2> gpurun_out/bench_${TAG}_reference_arm.err

# 2. configs[3]: persistent phase-synchronous kernel vs the one-launch kernel, launch list, ncu of the fit kernel
for fs in 1 0; do
  timeout 200 python bench.py --workload ransac --steps 10 --no-cpu-baseline --fit-stream $fs > gpurun_out/bench_${TAG}_ransac_fs$fs.json 2>/dev/null
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${TAG}_ransac.csv \
  python bench.py --workload ransac --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_pose_fit_stream -s 3 -c 1 -f -o gpurun_out/prof_fit_stream_$TAG \
  python bench.py --workload ransac --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1

# 3. LM parity table (default kernels with the LDL^T solve, exact-order mode, both builds of the reference)
timeout 300 python scripts/lm_parity_table.py --out gpurun_out/lm_parity_$TAG > /dev/null 2>&1

# 4. stage kernels after MATCH: ncu; stage chains of a batch: sweeps
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_ransac_first|k_ransac_refit|k_meanshift" -s 40 -c 6 -f -o gpurun_out/prof_stages_$TAG \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --other-configs 0 > /dev/null 2>&1
for cfg in "1 1 4" "1 0 4" "0 1 4" "1 1 1" "1 1 8"; do set -- $cfg
  timeout 300 python bench.py --other-configs 0 --steps 10 --no-cpu-baseline --batch-graph $1 --merge-levels $2 --pose-warps $3 > gpurun_out/bench_${TAG}_bg$1_ml$2_pw$3.json 2>/dev/null
done
timeout 120 ./scripts/probe/graph_branches.bin        # nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/probe/graph_branches.bin scripts/probe/graph_branches.cu

# 5. coarse kernel: database splits per query tile by shard size; ncu on a 125 k-row shard
DBG_OBJECTS=125 DBG_CONFIGS=1:0:1 DBG_SPLITS=0,2,3,4 DBG_ITERS=12 timeout 200 python scripts/gpu_coarse_dbg.py splits_125k >> gpurun_out/coarse_splits_$TAG.jsonl
DBG_OBJECTS=500 DBG_CONFIGS=1:0:1 DBG_SPLITS=0,3,4,6 DBG_ITERS=10 timeout 200 python scripts/gpu_coarse_dbg.py splits_500k >> gpurun_out/coarse_splits_$TAG.jsonl
DBG_CONFIGS=1:0:1 DBG_SPLITS=0,4,6 DBG_ITERS=8 timeout 300 python scripts/gpu_coarse_dbg.py splits_1m >> gpurun_out/coarse_splits_$TAG.jsonl
DBG_OBJECTS=125 DBG_CONFIGS=1:0:1 DBG_SPLITS=0 DBG_ITERS=4 timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_match_coarse -s 2 -c 1 -f \
  -o gpurun_out/prof_coarse_125k_$TAG python scripts/gpu_coarse_dbg.py ncu_125k > /dev/null 2>&1

# 6. N GPUs (gpurun --gpus N): the default line, and one frame sharded over the GPUs with cluster-distributed RANSAC
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus N --steps 20 --warmup 5
#   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus N --frames 1 --steps 20 --warmup 5
