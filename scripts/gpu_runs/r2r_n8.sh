#!/bin/bash
# round 2, call r (8 GPUs): sharded == single-GPU at world 8, the host-DMA ceiling with 8 ranks copying at once, then
# bench.py --gpus 8 exactly as the driver launches it (final code)
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29521 scripts/check_dist.py > gpurun_out/r2r_check_dist_n8.log 2>&1; echo "check_dist rc $?"; tail -22 gpurun_out/r2r_check_dist_n8.log | cut -c1-200
timeout 180 $TR --master-port 29522 scripts/probe_h2d.py > gpurun_out/r2r_probe_h2d_n8.log 2>&1; echo "h2d rc $?"; grep "copying" gpurun_out/r2r_probe_h2d_n8.log | cut -c1-250
NCCL_DEBUG=INFO timeout 1200 $TR --master-port 29523 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_r2r_n8.out 2> gpurun_out/bench_r2r_n8.err; echo "bench n8 rc $?"; tail -c 600 gpurun_out/bench_r2r_n8.err; tail -n 1 gpurun_out/bench_r2r_n8.out > gpurun_out/bench_r2r_n8.json; tail -n 1 gpurun_out/bench_r2r_n8.out | cut -c1-300
