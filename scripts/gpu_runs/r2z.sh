#!/bin/bash
# round 2, call z: final single-GPU validation of the last code — smoke, bench (as the driver runs it), whole GPU suite
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2z2_smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/r2z2_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2z2.out 2> gpurun_out/bench_r2z2.err; echo "bench rc $?"; tail -c 300 gpurun_out/bench_r2z2.err; tail -n 1 gpurun_out/bench_r2z2.out > gpurun_out/bench_r2z2.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2z2.json'))
e=d['e2e']; print(d['value'], d['ms_per_step'], d['cusolver_ms_per_step'], d['clocks']['nvml_call_ms_max'], d['roofline']['frac'])
print(e['value'], e['wall_s_each_step']); print(e['run_split_s_each_step'][-1]); print(d['cpu_baseline']['value'])
PY
timeout 1800 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/gputests_r2z2.log 2>&1; echo "gpu tests rc $?"; tail -14 gpurun_out/gputests_r2z2.log | cut -c1-250
