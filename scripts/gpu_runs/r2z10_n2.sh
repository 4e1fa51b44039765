#!/bin/bash
# round 2, call z10 (2 GPUs): headline under torchrun with the final defaults (subspace merge solve on every rank)
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --no-cpu --no-e2e > gpurun_out/bench_r2z10_n2.out 2> gpurun_out/bench_r2z10_n2.err; echo "bench n2 rc $?"; tail -c 300 gpurun_out/bench_r2z10_n2.err; tail -n 1 gpurun_out/bench_r2z10_n2.out > gpurun_out/bench_r2z10_n2.json; cut -c1-260 gpurun_out/bench_r2z10_n2.json
