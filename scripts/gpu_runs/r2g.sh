#!/bin/bash
# round 2, call g: whole GPU suite with the final defaults, smoke, bench (all sections), ncu of the fused sytrd
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2g_smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/r2g_smoke.log
timeout 200 python scripts/probe_eigh_cluster.py > gpurun_out/probe_eigh_r2g.log 2>&1; echo "probe rc $?"; head -3 gpurun_out/probe_eigh_r2g.log | cut -c1-250
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_launches_profile_kernels.csv python scripts/profile_kernels.py > gpurun_out/r2g_ncu_list.log 2>&1; echo "ncu list rc $?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; echo "bench rc $?"; tail -c 400 gpurun_out/bench_r2g.err
timeout 1800 python -m pytest tests -q -m gpu --durations=10 > gpurun_out/gputests_r2g.log 2>&1; echo "gpu tests rc $?"; tail -18 gpurun_out/gputests_r2g.log | cut -c1-250
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"sytrd_fused|backtransform" -c 2 -f -o gpurun_out/r2g_prof_eigh python scripts/profile_kernels.py > gpurun_out/r2g_ncu_eigh.log 2>&1; echo "ncu eigh rc $?"
