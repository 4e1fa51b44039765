#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_k1_gpu.py tests/test_reference_suite_gpu.py tests/test_edge_cases_gpu.py -q -m gpu > gpurun_out/pytest_k1.log 2>&1; echo "pytest k1+refsuite+edge rc $?" >> gpurun_out/summary.log
timeout 300 python scripts/bench_fullres.py > gpurun_out/fullres_bench.json 2> gpurun_out/fullres_bench.err; echo "bench fullres rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -6 gpurun_out/pytest_k1.log; python -c "
import json; d=json.load(open('gpurun_out/fullres_bench.json')); print(d['sbin3'], d['sbin7'])"; tail -3 gpurun_out/fullres_bench.err
