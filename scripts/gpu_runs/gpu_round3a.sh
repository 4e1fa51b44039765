#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_fullres_gpu.py -q -m gpu > gpurun_out/pytest_fullres.log 2>&1; echo "pytest fullres rc $?" >> gpurun_out/summary.log
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_fullres_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc $?" >> gpurun_out/summary.log
timeout 300 python scripts/bench_fullres.py > gpurun_out/fullres_bench.json 2> gpurun_out/fullres_bench.err; echo "bench fullres rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -40 gpurun_out/pytest_fullres.log; tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/fullres_bench.json; tail -3 gpurun_out/fullres_bench.err
