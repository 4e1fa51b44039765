#!/bin/bash
# round 2, call n: K-blocked projection operands (fm_retile_u8, fm_project_i8_blocked), configs[3] bench section
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 600 python -m pytest tests/test_project_i8_gpu.py tests/test_gemm_i8_gpu.py -q -m gpu -x > gpurun_out/r2n_tests.log 2>&1; echo "tests rc $?"; tail -6 gpurun_out/r2n_tests.log | cut -c1-300
timeout 300 python scripts/probe_blocked.py > gpurun_out/probe_blocked.log 2>&1; echo "probe rc $?"; cat gpurun_out/probe_blocked.log | tail -8
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r2n.json 2> gpurun_out/bench_r2n.err; echo "bench rc $?"; tail -c 300 gpurun_out/bench_r2n.err
