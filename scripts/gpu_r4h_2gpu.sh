#!/bin/bash
# 2 GPUs: one frame with the RANSAC tasks distributed by cluster (BASELINE configs[2] as a sharded frame), and the default line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus 2 --frames 1 --steps 20 --warmup 5 > gpurun_out/bench_r4h_cluster_2gpu.json 2> gpurun_out/bench_r4h_cluster_2gpu.err
echo "cluster partition 2 GPUs rc $?"; tail -2 gpurun_out/bench_r4h_cluster_2gpu.json | cut -c1-400; tail -3 gpurun_out/bench_r4h_cluster_2gpu.err
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r4h_2gpu.json 2> gpurun_out/bench_r4h_2gpu.err
echo "default 2 GPUs rc $?"; tail -1 gpurun_out/bench_r4h_2gpu.json | cut -c1-300
