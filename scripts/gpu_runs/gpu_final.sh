#!/bin/bash
# what the driver runs at round end: GPU tests, smoke, the bench line and the reference arm
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc $?" >> gpurun_out/summary.log
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc $?" >> gpurun_out/summary.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/summary.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc $?" >> gpurun_out/summary.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "bench reference rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -4 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; tail -2 gpurun_out/bench_default.err; cut -c1-1500 gpurun_out/bench_default.json; echo; cut -c1-600 gpurun_out/bench_reference.json
