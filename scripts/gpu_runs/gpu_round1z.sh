#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc $?" >> gpurun_out/summary.log
timeout 600 python scripts/probe_pipeline.py 0 0 > gpurun_out/pipe_impl0.log 2>&1; echo "pipe rc $?" >> gpurun_out/summary.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/summary.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc $?" >> gpurun_out/summary.log
timeout 300 python scripts/profile_kernels.py > gpurun_out/profile_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 2 -c 1 -f -o gpurun_out/prof_gemm python scripts/profile_kernels.py > gpurun_out/ncu_gemm.log 2>&1
echo "ncu full rc $?" >> gpurun_out/summary.log
timeout 900 python bench.py --steps 1 --warmup 5 --no-e2e --no-cpu > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err && \
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bin_motion|gemm_|splitk|gram_assemble|row_means|row_dot|topk|transpose_kernel|mean_update|gather_roi|split_lo|sign_from" -s 1150 -c 240 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 5 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; grep "^FAILED" gpurun_out/pytest_gpu.log | cut -c1-200; tail -2 gpurun_out/pytest_gpu.log | cut -c1-300; grep -E "S rel err|orth|trace" gpurun_out/pipe_impl0.log | cut -c1-250; tail -1 gpurun_out/smoke.log
for f in gpurun_out/bench_n1; do grep -o '"value": [0-9.]*' $f.json | head -1; grep -o '"ms_per_step": [0-9.]*' $f.json | head -1; grep -o '"host_ms_per_step[^]]*]' $f.json; grep -o '"gpu_launches": [0-9]*' $f.json; grep -o '"e2e": {[^}]*' $f.json | cut -c1-300; grep -o '"stage_ms": {[^}]*}' $f.json | tail -1 | cut -c1-1400; done; wc -l gpurun_out/launches.csv
