#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 180 python scripts/probe_gemm.py 9 > gpurun_out/gemm_impl9.log 2>&1; echo "gemm9 rc $?" >> gpurun_out/summary.log
timeout 180 python scripts/probe_gemm.py 9 presplit > gpurun_out/gemm_impl9_b.log 2>&1; echo "gemm9 presplit rc $?" >> gpurun_out/summary.log
timeout 180 python scripts/probe_gemm.py 9 presplit both > gpurun_out/gemm_impl9_both.log 2>&1; echo "gemm9 both rc $?" >> gpurun_out/summary.log
timeout 180 python scripts/probe_gemm.py 0 presplit > gpurun_out/gemm_impl0_b.log 2>&1
timeout 900 python -m pytest tests/test_gemm_gpu.py -q > gpurun_out/pytest_gemm.log 2>&1; echo "pytest gemm rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -14 gpurun_out/gemm_impl9.log; tail -6 gpurun_out/gemm_impl9_b.log; tail -6 gpurun_out/gemm_impl9_both.log; tail -6 gpurun_out/gemm_impl0_b.log; tail -8 gpurun_out/pytest_gemm.log | cut -c1-300
