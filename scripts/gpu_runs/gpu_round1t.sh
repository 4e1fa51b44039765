#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python scripts/sweep.py > gpurun_out/sweep.log 2>&1; echo "sweep rc $?"
cat gpurun_out/sweep.log | cut -c1-300
