#!/bin/bash
# round 2, call x: proc file written while the pass runs (run / save tests, bench line with the run() split)
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_run_config.py tests/test_save_parity.py tests/test_reference_suite_gpu.py tests/test_video_files.py tests/test_edge_cases_gpu.py -x -q -m gpu > gpurun_out/r2x_tests.log 2>&1; echo "tests rc $?"; tail -15 gpurun_out/r2x_tests.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 5 --e2e-steps 5 --no-extras --no-cpu > gpurun_out/bench_r2x.out 2> gpurun_out/bench_r2x.err; echo "bench rc $?"; tail -c 300 gpurun_out/bench_r2x.err; tail -n 1 gpurun_out/bench_r2x.out > gpurun_out/bench_r2x.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r2x.json'))
print(d['value'], d['ms_per_step'], d['clocks'])
e=d['e2e']; print({k:e[k] for k in e if k not in ('stage_ms','timeline_ms','api')})
PY
