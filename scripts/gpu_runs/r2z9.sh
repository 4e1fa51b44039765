#!/bin/bash
# round 2, call z9: bench line with the round's final defaults (merge eigensolve by subspace iteration)
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 170 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu --e2e-steps 3 > gpurun_out/bench_r2z9.out 2> gpurun_out/bench_r2z9.err; echo "bench rc $?"; tail -c 300 gpurun_out/bench_r2z9.err; tail -n 1 gpurun_out/bench_r2z9.out > gpurun_out/bench_r2z9.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2z9.json'))
print(d['value'], d['ms_per_step'], d['cusolver_ms_per_step'], d['roofline']['frac'])
print([round(x,1) for x in d['host_ms_per_step']])
print(d['e2e']['value'], d['e2e']['wall_s_each_step'])
PY
