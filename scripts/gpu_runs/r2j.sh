#!/bin/bash
# round 2, call j: cuSOLVER sytrd split; bench with main-thread clock sampling
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 200 python scripts/probe_sytrd.py > gpurun_out/probe_sytrd.log 2>&1; echo "probe sytrd rc $?"; cat gpurun_out/probe_sytrd.log | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-extras --no-e2e > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err; echo "bench rc $?"; tail -c 300 gpurun_out/bench_r2j.err
