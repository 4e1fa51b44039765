#!/bin/bash
# multi-GPU scaling: bench at N=4 and N=8 (and the dist check at 8)
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
free -g > gpurun_out/env8.log; nproc >> gpurun_out/env8.log; nvidia-smi topo -m >> gpurun_out/env8.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/check_dist.py > gpurun_out/check_dist8.log 2>&1; echo "check_dist8 rc $?" >> gpurun_out/summary.log
for N in 8 4; do
FM_BENCH_DIAG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc $?" >> gpurun_out/summary.log
done
cat gpurun_out/summary.log; head -3 gpurun_out/env8.log; tail -3 gpurun_out/check_dist8.log
for N in 8 4; do f=gpurun_out/bench_n$N; grep DIAG $f.err | head -2 | cut -c1-300; grep -o '"value": [0-9.]*' $f.json | head -1; grep -o '"ms_per_step": [0-9.]*' $f.json | head -1; grep -o '"host_ms_per_step[^]]*]' $f.json; grep -o '"e2e": {[^}]*' $f.json | cut -c1-400; tail -3 $f.err | cut -c1-300; done
