#!/bin/bash
# round 2, call o: fused sbin-1 stage-3 pass (fm_motion_ref_planes) against the kernels it replaces, goldens at sbin 1
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests/test_k1_gpu.py tests/test_project_i8_gpu.py -q -m gpu -x > gpurun_out/r2o_tests.log 2>&1; echo "tests rc $?"; tail -6 gpurun_out/r2o_tests.log | cut -c1-300
timeout 600 python -m pytest tests/test_large_golden_gpu.py -q -m gpu -k "c4_crop or c1_s4" -s > gpurun_out/r2o_golden.log 2>&1; echo "golden rc $?"; grep -E "passed|failed|S rel" gpurun_out/r2o_golden.log | cut -c1-200
timeout 300 python scripts/probe_c3.py motion > gpurun_out/probe_c3_motion.log 2>&1; echo "probe rc $?"; tail -4 gpurun_out/probe_c3_motion.log | cut -c1-250
