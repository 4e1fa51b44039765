#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python scripts/bench_fullres.py > gpurun_out/fullres_bench.json 2> gpurun_out/fullres_bench.err; echo "bench fullres rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; cat gpurun_out/fullres_bench.json; tail -5 gpurun_out/fullres_bench.err
