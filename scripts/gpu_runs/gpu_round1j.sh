#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc $?" >> gpurun_out/summary.log
timeout 900 python scripts/scale_check.py c3 > gpurun_out/scale_c3.log 2>&1; echo "c3 rc $?" >> gpurun_out/summary.log
timeout 900 python scripts/scale_check.py c4 > gpurun_out/scale_c4.log 2>&1; echo "c4 rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300; tail -12 gpurun_out/scale_c3.log | cut -c1-600; tail -10 gpurun_out/scale_c4.log | cut -c1-600
