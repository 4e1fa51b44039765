#!/bin/bash
# round 2, call l: register-accumulator tridiagonalisation (FM_SYTRD_IMPL=2) against the shared-memory one, blocked
# Gram / left-vector test, e2e with the parallel .npy writer
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
FM_SYTRD_IMPL=2 timeout 300 python -m pytest tests/test_eigh_cluster_gpu.py -q -m gpu -x > gpurun_out/r2l_eigh_impl2.log 2>&1; echo "eigh tests impl 2 rc $?"; tail -4 gpurun_out/r2l_eigh_impl2.log | cut -c1-250
FM_SYTRD_IMPL=2 timeout 200 python scripts/probe_eigh_cluster.py > gpurun_out/probe_eigh_impl2.log 2>&1; echo "probe impl 2 rc $?"; head -4 gpurun_out/probe_eigh_impl2.log | cut -c1-250
timeout 200 python scripts/probe_eigh_cluster.py > gpurun_out/probe_eigh_impl1.log 2>&1; echo "probe impl 1 rc $?"; head -4 gpurun_out/probe_eigh_impl1.log | cut -c1-250
timeout 300 python -m pytest tests/test_gemm_i8_gpu.py -q -m gpu > gpurun_out/r2l_i8_tests.log 2>&1; echo "i8 tests rc $?"; tail -4 gpurun_out/r2l_i8_tests.log | cut -c1-250
FM_SYTRD_IMPL=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_r2l.json 2> gpurun_out/bench_r2l.err; echo "bench rc $?"; tail -c 300 gpurun_out/bench_r2l.err
