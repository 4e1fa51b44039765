#!/bin/bash
# round 2, call v (4 GPUs, never run before this round): sharded == single GPU at world 4, then bench.py --gpus 4 as the driver launches it
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29531 scripts/check_dist.py > gpurun_out/r2v_check_dist_n4.log 2>&1; echo "check_dist rc $?"; tail -3 gpurun_out/r2v_check_dist_n4.log | cut -c1-200
NCCL_DEBUG=INFO timeout 1200 $TR --master-port 29533 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_r2v_n4.out 2> gpurun_out/bench_r2v_n4.err; echo "bench n4 rc $?"; tail -c 300 gpurun_out/bench_r2v_n4.err; tail -n 1 gpurun_out/bench_r2v_n4.out > gpurun_out/bench_r2v_n4.json; tail -n 1 gpurun_out/bench_r2v_n4.out | cut -c1-300
