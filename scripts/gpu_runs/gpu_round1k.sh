#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 180 python scripts/probe_gemm.py 0 presplit > gpurun_out/gemm_presplit_b.log 2>&1
timeout 180 python scripts/probe_gemm.py 0 presplit both > gpurun_out/gemm_presplit_both.log 2>&1; echo "gemm both rc $?" >> gpurun_out/summary.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc $?" >> gpurun_out/summary.log
FM_BENCH_DIAG=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300; tail -6 gpurun_out/gemm_presplit_b.log; tail -6 gpurun_out/gemm_presplit_both.log
for f in gpurun_out/bench_n1; do grep DIAG $f.err; grep -o '"value": [0-9.]*' $f.json | head -1; grep -o '"ms_per_step": [0-9.]*' $f.json | head -1; grep -o '"host_ms_per_step[^]]*]' $f.json; grep -o '"stage_ms": {[^}]*}' $f.json | tail -1 | cut -c1-1500; done
