#!/bin/bash
# round 2, call y: e2e with a fresh output folder per call; where the tail of run() goes
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 600 python bench.py --steps 5 --warmup 5 --e2e-steps 6 --no-extras --no-cpu > gpurun_out/bench_r2y2.out 2> gpurun_out/bench_r2y2.err; echo "bench rc $?"; tail -n 1 gpurun_out/bench_r2y2.out > gpurun_out/bench_r2y2.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2y2.json'))
e=d['e2e']; print(d['ms_per_step'], e['value'], e['wall_s_each_step'])
for s in e['run_split_s_each_step']: print(s)
PY
