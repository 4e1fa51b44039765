#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
N=${1:-2}
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc $?" >> gpurun_out/summary.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc $?" >> gpurun_out/summary.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/check_dist.py > gpurun_out/check_dist.log 2>&1; echo "check_dist rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -2 gpurun_out/check_dist.log
for f in gpurun_out/bench_n1 gpurun_out/bench_n$N; do grep -o '"value": [0-9.]*' $f.json | head -1; grep -o '"ms_per_step": [0-9.]*' $f.json | head -1; grep -o '"host_ms_per_step[^]]*]' $f.json; grep -o '"cusolver_ms_each_step[^]]*]' $f.json; grep -o '"e2e": {[^}]*' $f.json | cut -c1-330; done
