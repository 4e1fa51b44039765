#!/bin/bash
# round 2, call u: where run()'s wall time goes (pass / close / save split), sparser NVML sampling
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 900 python bench.py --steps 20 --warmup 5 --e2e-steps 5 --no-extras --no-cpu > gpurun_out/bench_r2u.out 2> gpurun_out/bench_r2u.err; echo "bench rc $?"; tail -c 300 gpurun_out/bench_r2u.err; tail -n 1 gpurun_out/bench_r2u.out > gpurun_out/bench_r2u.json; cut -c1-200 gpurun_out/bench_r2u.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2u.json'))
print(d['ms_per_step'], d['clocks'])
e=d['e2e']; print({k:e[k] for k in e if k not in ('stage_ms','timeline_ms','api')})
PY
df -h /tmp | tail -2; mount | grep -E " /tmp | / " | head -3
