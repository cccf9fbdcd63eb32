#!/bin/bash
# 8 GPUs: the default line (pipelined schedule chosen automatically) and one frame sharded over 8 GPUs with cluster-distributed RANSAC
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r4i_8gpu.json 2> gpurun_out/bench_r4i_8gpu.err
echo "default 8 GPUs rc $?"; head -c 400 gpurun_out/bench_r4i_8gpu.json; echo
timeout 200 $TR --master-port 29522 bench.py --gpus 8 --frames 1 --steps 20 --warmup 5 > gpurun_out/bench_r4i_cluster_8gpu.json 2> gpurun_out/bench_r4i_cluster_8gpu.err
echo "cluster partition 8 GPUs rc $?"; head -c 300 gpurun_out/bench_r4i_cluster_8gpu.json; echo; tail -2 gpurun_out/bench_r4i_cluster_8gpu.err
