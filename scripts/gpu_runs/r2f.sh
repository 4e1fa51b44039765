#!/bin/bash
# round 2, call f: one-sweep tridiagonalisation + faster bisection / inverse iteration / back-transformation,
# configs[3] probes (projection at K = 1.31 M, fm_motion_ref fast path), bench
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 300 python -m pytest tests/test_eigh_cluster_gpu.py tests/test_k1_gpu.py -q -m gpu -x > gpurun_out/r2f_tests.log 2>&1; echo "tests rc $?"; tail -6 gpurun_out/r2f_tests.log | cut -c1-300
timeout 200 python scripts/probe_eigh_cluster.py > gpurun_out/probe_eigh_fused.log 2>&1; echo "probe fused rc $?"; head -3 gpurun_out/probe_eigh_fused.log | cut -c1-250
FM_SYTRD_IMPL=0 timeout 200 python scripts/probe_eigh_cluster.py > gpurun_out/probe_eigh_twosweep.log 2>&1; echo "probe two-sweep rc $?"; head -2 gpurun_out/probe_eigh_twosweep.log | cut -c1-250
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2f_launches_profile_kernels.csv python scripts/profile_kernels.py > gpurun_out/r2f_ncu_list.log 2>&1; echo "ncu list rc $?"
timeout 400 python scripts/probe_c3.py all > gpurun_out/probe_c3.log 2>&1; echo "probe c3 rc $?"; tail -14 gpurun_out/probe_c3.log | cut -c1-250
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; echo "bench rc $?"; tail -c 600 gpurun_out/bench_r2f.err
timeout 400 ncu --set full --clock-control none --import-source on -k regex:project_i8 -c 1 -f -o gpurun_out/r2f_prof_proj_c3 python scripts/probe_c3.py proj > gpurun_out/r2f_ncu_proj.log 2>&1; echo "ncu proj c3 rc $?"
ncu -i gpurun_out/r2f_prof_proj_c3.ncu-rep --page raw --csv > gpurun_out/r2f_prof_proj_c3_raw.csv 2>/dev/null
