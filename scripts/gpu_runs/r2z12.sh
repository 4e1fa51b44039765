#!/bin/bash
# round 2, call z12: short bench line on the round's last commit
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 60 python bench.py --steps 10 --warmup 5 --no-extras --no-cpu --no-e2e > gpurun_out/bench_r2z12.out 2> gpurun_out/bench_r2z12.err; echo "bench rc $?"; tail -c 200 gpurun_out/bench_r2z12.err; tail -n 1 gpurun_out/bench_r2z12.out > gpurun_out/bench_r2z12.json; cut -c1-230 gpurun_out/bench_r2z12.json
