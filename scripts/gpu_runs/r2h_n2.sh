#!/bin/bash
# round 2 (2 GPUs): bench.py --gpus 2 exactly as the driver launches it, on the round's last code
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2z5_n2.out 2> gpurun_out/bench_r2z5_n2.err; echo "bench n2 rc $?"; tail -c 300 gpurun_out/bench_r2z5_n2.err; tail -n 1 gpurun_out/bench_r2z5_n2.out > gpurun_out/bench_r2z5_n2.json; cut -c1-300 gpurun_out/bench_r2z5_n2.json
