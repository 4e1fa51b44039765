#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_dist.py > gpurun_out/check_dist.log 2>&1; echo "check_dist rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -25 gpurun_out/check_dist.log
