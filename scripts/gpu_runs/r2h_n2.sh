#!/bin/bash
# round 2, call h (2 GPUs): the sharded path against the single-GPU path, then bench.py --gpus 2 exactly as the driver launches it
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_dist.py > gpurun_out/r2h_check_dist_n2.log 2>&1; echo "check_dist rc $?"; tail -22 gpurun_out/r2h_check_dist_n2.log | cut -c1-200
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r2h_n2.out 2> gpurun_out/bench_r2h_n2.err; echo "bench n2 rc $?"; tail -c 600 gpurun_out/bench_r2h_n2.err; tail -n 1 gpurun_out/bench_r2h_n2.out > gpurun_out/bench_r2h_n2.json; grep -c NCCL gpurun_out/bench_r2h_n2.out
