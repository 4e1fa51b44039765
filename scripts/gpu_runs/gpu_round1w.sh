#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc $?" >> gpurun_out/summary.log
timeout 600 python scripts/probe_pipeline.py 0 0 > gpurun_out/pipe_impl0.log 2>&1; echo "pipe rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; grep "^FAILED" gpurun_out/pytest_gpu.log | cut -c1-200; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300; grep -E "S rel err|orth|trace" gpurun_out/pipe_impl0.log | cut -c1-250
