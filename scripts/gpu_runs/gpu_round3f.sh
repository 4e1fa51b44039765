#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_reference_suite_gpu.py -q -m gpu > gpurun_out/pytest_refsuite.log 2>&1; echo "pytest refsuite rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -30 gpurun_out/pytest_refsuite.log
