#!/bin/bash
# round 2, call k (8 GPUs): bench.py --gpus 8 exactly as the driver launches it
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
NCCL_DEBUG=INFO timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r2k_n8.out 2> gpurun_out/bench_r2k_n8.err; echo "bench n8 rc $?"; tail -c 600 gpurun_out/bench_r2k_n8.err; tail -n 1 gpurun_out/bench_r2k_n8.out > gpurun_out/bench_r2k_n8.json; tail -n 1 gpurun_out/bench_r2k_n8.out | cut -c1-300
