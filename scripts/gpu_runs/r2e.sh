#!/bin/bash
# round 2, call e: per-kernel times of the hand-written eigensolver, fixed tests, bench with the bin pass moved under
# the merge solve, configs[3] memory fixes, ncu --set full of the default stage-3 kernels
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 300 python -m pytest tests/test_video_files.py tests/test_edge_cases_gpu.py tests/test_eigh_cluster_gpu.py -q -m gpu > gpurun_out/r2e_tests.log 2>&1; echo "tests rc $?"; tail -6 gpurun_out/r2e_tests.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2e.json 2> gpurun_out/bench_r2e.err; echo "bench rc $?"; tail -c 1200 gpurun_out/bench_r2e.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_launches_profile_kernels.csv python scripts/profile_kernels.py > gpurun_out/r2e_ncu_list.log 2>&1; echo "ncu list rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bin_motion|project_i8|gemm_i8|sytrd_cluster" -f -o gpurun_out/r2e_prof python scripts/profile_kernels.py > gpurun_out/r2e_ncu_full.log 2>&1; echo "ncu full rc $?"
ncu -i gpurun_out/r2e_prof.ncu-rep --page raw --csv > gpurun_out/r2e_prof_raw.csv 2>/dev/null
