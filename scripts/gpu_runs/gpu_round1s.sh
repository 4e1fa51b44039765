#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log
for f in gpurun_out/bench_n8; do grep -o '"value": [0-9.]*' $f.json | head -1; grep -o '"ms_per_step": [0-9.]*' $f.json | head -1; grep -o '"host_ms_per_step[^]]*]' $f.json; grep -o '"e2e": {.*"timeline_ms": {[^}]*}' $f.json | cut -c1-1500; grep -o '"stage_ms": {[^}]*}' $f.json | tail -1 | cut -c1-1200; done; tail -3 gpurun_out/bench_n8.err | cut -c1-300
