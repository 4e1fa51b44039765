#!/bin/bash
# round 2, call w (4 GPUs): configs[2] with the merges dealt by a cost linear in the problem order; per-rank times
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29543 bench.py --gpus 4 --steps 5 --warmup 5 --no-e2e --c3-frames 0 > gpurun_out/bench_r2w_n4.out 2> gpurun_out/bench_r2w_n4.err; echo "bench n4 rc $?"; tail -c 300 gpurun_out/bench_r2w_n4.err; tail -n 1 gpurun_out/bench_r2w_n4.out > gpurun_out/bench_r2w_n4.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2w_n4.json'))
print(d['value'], d['ms_per_step'])
c=d['extras']['configs2']; print(c['ms_per_step']); print(json.dumps(c['stage_ms_rank0']))
for r in c.get('per_rank',[]): print(r)
PY
