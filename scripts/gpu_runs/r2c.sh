#!/bin/bash
# round 2, call c: widened gates (every singular value, every gap-resolved component), the decode-cache seam checks,
# and the rewritten bench.py (reference-arm model, process.run e2e, side measurements) at N = 1
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc $?"; tail -c 1500 gpurun_out/bench_r2c.err; tail -c 6000 gpurun_out/bench_r2c.json | cut -c1-6000
timeout 1500 python -m pytest tests -q -m gpu -x --durations=8 -k "not large_case" > gpurun_out/gputests_r2c.log 2>&1; echo "gpu tests rc $?"; tail -25 gpurun_out/gputests_r2c.log | cut -c1-300
timeout 900 python -m pytest tests/test_large_golden_gpu.py tests/test_all_components_gpu.py -q -m gpu -s > gpurun_out/gputests_r2c_large.log 2>&1; echo "large tests rc $?"; grep -E "resolved|S rel|passed|failed|Error" gpurun_out/gputests_r2c_large.log | cut -c1-250 | tail -50
timeout 300 python scripts/probe_accuracy.py i8b > gpurun_out/accuracy_i8b.log 2>&1; echo "probe_accuracy rc $?"
