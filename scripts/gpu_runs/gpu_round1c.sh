#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc $?" >> gpurun_out/summary.log
timeout 300 python scripts/probe_trunc.py > gpurun_out/trunc.log 2>&1; echo "trunc rc $?" >> gpurun_out/summary.log
for impl in 0 5 3 6; do timeout 300 python scripts/probe_gemm.py $impl > gpurun_out/gemm_impl$impl.log 2>&1; done
timeout 600 python scripts/probe_eig.py > gpurun_out/eig.log 2>&1; echo "eig rc $?" >> gpurun_out/summary.log
timeout 600 python scripts/probe_pipeline.py 0 0 > gpurun_out/pipe_impl0.log 2>&1; echo "pipe rc $?" >> gpurun_out/summary.log
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -25 gpurun_out/pytest_gpu.log | cut -c1-200; cat gpurun_out/trunc.log; tail -5 gpurun_out/gemm_impl5.log; tail -4 gpurun_out/eig.log; cut -c1-700 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
