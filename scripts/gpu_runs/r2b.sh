#!/bin/bash
# round 2, call b: eigensolve concurrency probe, all-component accuracy profile with the new defaults (i8 Gram at both
# levels, i8 projection), a short bench, then the whole GPU suite without opt-in gates
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 400 python scripts/probe_eig.py > gpurun_out/probe_eig.log 2>&1; echo "probe_eig rc $?"; tail -30 gpurun_out/probe_eig.log
timeout 600 python scripts/probe_accuracy.py i8 > gpurun_out/accuracy_i8.log 2>&1; echo "probe_accuracy rc $?"; tail -60 gpurun_out/accuracy_i8.log | cut -c1-200
FM_GRAM_MODE=tf32x3 FM_PROJ_MODE=tf32x3 timeout 600 python scripts/probe_accuracy.py tf32x3 > gpurun_out/accuracy_tf32x3.log 2>&1; echo "probe_accuracy tf32 rc $?"
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc $?"; tail -c 3000 gpurun_out/bench_r2b.json
timeout 1500 python -m pytest tests -q -m gpu -x --durations=15 > gpurun_out/gputests_r2b.log 2>&1; echo "gpu tests rc $?"; tail -40 gpurun_out/gputests_r2b.log | cut -c1-250
