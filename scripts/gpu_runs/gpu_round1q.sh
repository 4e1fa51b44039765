#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc $?" >> gpurun_out/summary.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc $?" >> gpurun_out/summary.log
timeout 300 python scripts/sanitize_small.py > gpurun_out/sanitize_plain.log 2>&1 && \
  timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/sanitize_small.py > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300; tail -6 gpurun_out/sanitize_memcheck.log | cut -c1-300
for f in gpurun_out/bench_n1; do grep -o '"value": [0-9.]*' $f.json | head -1; grep -o '"ms_per_step": [0-9.]*' $f.json | head -1; grep -o '"host_ms_per_step[^]]*]' $f.json; grep -o '"e2e": {[^}]*' $f.json | cut -c1-300; done
