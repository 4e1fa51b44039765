#!/bin/bash
# round 2, call q: ncu --set full of the K-blocked projection at K = 1 310 720 and of the fused sbin-1 stage-3 pass; edge tests
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 300 python -m pytest tests/test_edge_cases_gpu.py tests/test_k1_gpu.py -q -m gpu > gpurun_out/r2q_tests.log 2>&1; echo "tests rc $?"; tail -3 gpurun_out/r2q_tests.log | cut -c1-200
timeout 500 ncu --set full --clock-control none --import-source on -k regex:project_i8 -s 4 -c 1 -f -o gpurun_out/r2q_prof_proj_blocked python scripts/probe_blocked.py > gpurun_out/r2q_ncu.log 2>&1; echo "ncu rc $?"
ncu -i gpurun_out/r2q_prof_proj_blocked.ncu-rep --page raw --csv > gpurun_out/r2q_prof_project_i8_blocked_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:motion_ref_u8 -s 6 -c 1 -f -o gpurun_out/r2q_prof_mref python scripts/probe_c3.py motion > gpurun_out/r2q_ncu2.log 2>&1; echo "ncu2 rc $?"
ncu -i gpurun_out/r2q_prof_mref.ncu-rep --page raw --csv > gpurun_out/r2q_prof_motion_ref_planes_raw.csv 2>/dev/null
