#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_reference_suite_gpu.py tests/test_k1_gpu.py -q -m gpu > gpurun_out/pytest_refsuite.log 2>&1; echo "pytest refsuite+k1 rc $?" >> gpurun_out/summary.log
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_reference_suite_gpu.py --deselect tests/test_k1_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -30 gpurun_out/pytest_refsuite.log; tail -5 gpurun_out/pytest_gpu.log
