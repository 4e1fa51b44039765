#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc $?" >> gpurun_out/summary.log
timeout 600 python scripts/probe_pipeline.py 0 0 > gpurun_out/pipe_impl0.log 2>&1; echo "pipe rc $?" >> gpurun_out/summary.log
FM_BENCH_DIAG=1 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -5 gpurun_out/pytest_gpu.log | cut -c1-200; grep -E "S rel err|orth|trace" gpurun_out/pipe_impl0.log | cut -c1-250
for f in gpurun_out/bench_n1; do grep DIAG $f.err; grep -o '"value": [0-9.]*' $f.json | head -1; grep -o '"ms_per_step": [0-9.]*' $f.json | head -1; grep -o '"host_ms_per_step[^]]*]' $f.json; grep -o '"e2e": {.*"stage_ms": {[^}]*}}' $f.json | cut -c1-900; grep -o '"roofline": {.*"cpu_baseline": {[^}]*}' $f.json | cut -c1-1800; done
