#!/bin/bash
# round 2, call p: final single-GPU validation — smoke, whole GPU suite, probes of the configs[3] stage-3 pass, bench
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2p_smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/r2p_smoke.log
timeout 300 python scripts/probe_c3.py motion > gpurun_out/probe_c3_motion.log 2>&1; echo "probe rc $?"; tail -6 gpurun_out/probe_c3_motion.log | cut -c1-250
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2p.json 2> gpurun_out/bench_r2p.err; echo "bench rc $?"; tail -c 300 gpurun_out/bench_r2p.err
timeout 1800 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/gputests_r2p.log 2>&1; echo "gpu tests rc $?"; tail -14 gpurun_out/gputests_r2p.log | cut -c1-250
