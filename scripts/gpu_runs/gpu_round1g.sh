#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 180 python scripts/probe_gemm.py 0 > gpurun_out/gemm_impl0.log 2>&1
timeout 180 python scripts/probe_gemm.py 0 presplit > gpurun_out/gemm_impl0_presplit.log 2>&1; echo "gemm presplit rc $?" >> gpurun_out/summary.log
timeout 600 python scripts/probe_eig.py > gpurun_out/eig.log 2>&1; echo "eig rc $?" >> gpurun_out/summary.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc $?" >> gpurun_out/summary.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc $?" >> gpurun_out/summary.log
timeout 300 python scripts/profile_kernels.py > gpurun_out/profile_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 2 -c 1 -f -o gpurun_out/prof_gemm_presplit python scripts/profile_kernels.py > gpurun_out/ncu_gemm.log 2>&1
echo "ncu rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -5 gpurun_out/pytest_gpu.log | cut -c1-200; tail -6 gpurun_out/gemm_impl0.log; tail -6 gpurun_out/gemm_impl0_presplit.log; tail -8 gpurun_out/eig.log; cut -c1-300 gpurun_out/bench_n1.json; grep -o '"host_ms_per_step[^]]*]' gpurun_out/bench_n1.json; grep -o '"e2e": {[^}]*}' gpurun_out/bench_n1.json; grep -o '"stage3[^,]*' gpurun_out/bench_n1.json
