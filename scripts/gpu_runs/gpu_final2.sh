#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc $?" >> gpurun_out/summary.log
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc $?" >> gpurun_out/summary.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log
