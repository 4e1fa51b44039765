#!/bin/bash
# round 2, call z3: per-pass driver calls (stream creation, cudaMemGetInfo) taken out of the step — stage-3 / pipeline tests, two bench lines
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_edge_cases_gpu.py tests/test_video_files.py tests/test_fullres_gpu.py tests/test_reference_suite_gpu.py tests/test_project_i8_gpu.py tests/test_eigh_cluster_gpu.py -x -q -m gpu > gpurun_out/r2z3_tests.log 2>&1; echo "tests rc $?"; tail -4 gpurun_out/r2z3_tests.log | cut -c1-300
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu > gpurun_out/bench_r2z3_$i.out 2> gpurun_out/bench_r2z3_$i.err; echo "bench rc $?"; tail -n 1 gpurun_out/bench_r2z3_$i.out > gpurun_out/bench_r2z3_$i.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2z3_$i.json'))
print(d['value'], d['ms_per_step'], d['clocks']['nvml_call_ms_max'])
print([round(x,1) for x in d['host_ms_per_step']])
print(d['e2e']['value'], d['e2e']['wall_s_each_step'])
PY
done
