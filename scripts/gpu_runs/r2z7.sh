#!/bin/bash
# round 2, call z7: the merge eigensolve by block subspace iteration (FM_MERGE_EIG=subspace): bench line, then the tests
# whose merge problem is large enough to take that path (20 000- and 100 000-frame cases)
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
export FM_MERGE_EIG=subspace
timeout 150 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu --no-e2e > gpurun_out/bench_r2z7.out 2> gpurun_out/bench_r2z7.err; echo "bench rc $?"; tail -c 300 gpurun_out/bench_r2z7.err; tail -n 1 gpurun_out/bench_r2z7.out > gpurun_out/bench_r2z7.json
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2z7.json'))
print(d['value'], d['ms_per_step'], d['cusolver_ms_per_step'])
print([round(x,1) for x in d['host_ms_per_step']])
print({k:v for k,v in d['stage_ms'].items() if 'merge' in k})
PY
timeout 280 python -m pytest tests/test_large_golden_gpu.py -x -q -m gpu -k c2_20k_s4 > gpurun_out/r2z7_golden.log 2>&1; echo "golden rc $?"; tail -3 gpurun_out/r2z7_golden.log | cut -c1-300
timeout 150 python -m pytest tests/test_fullsize_gpu.py -x -q -m gpu > gpurun_out/r2z7_fullsize.log 2>&1; echo "fullsize rc $?"; tail -3 gpurun_out/r2z7_fullsize.log | cut -c1-300
