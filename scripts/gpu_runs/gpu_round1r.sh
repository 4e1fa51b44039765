#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
N=${1:-2}
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc $?" >> gpurun_out/summary.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/summary.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc $?" >> gpurun_out/summary.log
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc $?" >> gpurun_out/summary.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300; tail -1 gpurun_out/smoke.log
for f in gpurun_out/bench_n1 gpurun_out/bench_n$N; do grep -o '"value": [0-9.]*' $f.json | head -1; grep -o '"ms_per_step": [0-9.]*' $f.json | head -1; grep -o '"host_ms_per_step[^]]*]' $f.json; grep -o '"e2e": {.*"timeline_ms": {[^}]*}' $f.json | cut -c1-1200; done; cut -c1-400 gpurun_out/bench_ref.json; head -c 300 gpurun_out/bench_n$N.json
