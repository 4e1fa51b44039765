#!/bin/bash
# round 2, call i: where cuSOLVER's merge solve spends its time; bench with the quieter clock sampler
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 200 python scripts/probe_sytrd.py > gpurun_out/probe_sytrd.log 2>&1; echo "probe sytrd rc $?"; cat gpurun_out/probe_sytrd.log | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; echo "bench rc $?"; tail -c 300 gpurun_out/bench_r2i.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r2i_ref.json 2> gpurun_out/bench_r2i_ref.err; echo "ref arm rc $?"; tail -3 gpurun_out/bench_r2i_ref.err
