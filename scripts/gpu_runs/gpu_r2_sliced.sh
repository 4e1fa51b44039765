#!/bin/bash
# Round-2 opener: validate the sliced Gram path (written without a GPU at the end of round 1) and time it.
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
FM_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_sliced_gram_gpu.py -x -q -m gpu -s > gpurun_out/sliced_tests.log 2>&1; echo "sliced tests rc $?"
tail -15 gpurun_out/sliced_tests.log
FM_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gemm_i8_gpu.py -q -m gpu -s > gpurun_out/i8_tests.log 2>&1; echo "i8 gemm tests rc $?"
tail -5 gpurun_out/i8_tests.log
FM_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_large_golden_gpu.py -x -q -m gpu -s > gpurun_out/large_golden_tests.log 2>&1; echo "large golden tests rc $?"
grep -E "S rel err|kres|passed|failed" gpurun_out/large_golden_tests.log | tail -8
FM_EXPERIMENTAL=1 FM_TEST_I8_IMPLS=0 timeout 600 python -m pytest tests/test_project_i8_gpu.py -q -m gpu -s > gpurun_out/project_i8_tests.log 2>&1; echo "i8 projection tests (one CTA per tile) rc $?"
FM_EXPERIMENTAL=1 FM_TEST_I8_IMPLS=1 timeout 600 python -m pytest tests/test_project_i8_gpu.py -q -m gpu -s -k "not planes_exact and not motion_rois" > gpurun_out/project_i8_pair_tests.log 2>&1; echo "i8 projection tests (CTA pairs) rc $?"
grep -E "passed|failed|Error|timeout" gpurun_out/project_i8_pair_tests.log | tail -4
grep -E "trace difference|passed|failed|Error|timeout" gpurun_out/project_i8_tests.log | tail -12
timeout 300 python scripts/probe_project_i8.py > gpurun_out/probe_project_i8.log 2>&1; echo "probe i8 projection rc $?"; tail -4 gpurun_out/probe_project_i8.log
FM_PROJ_MODE=i8 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_proj_i8.json 2> gpurun_out/bench_proj_i8.err; echo "bench FM_PROJ_MODE=i8 rc $?"; tail -c 1500 gpurun_out/bench_proj_i8.json
FM_PROJ_MODE=i8 FM_PROJ_I8_IMPL=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_proj_i8_pair.json 2> gpurun_out/bench_proj_i8_pair.err; echo "bench FM_PROJ_MODE=i8 pair rc $?"; tail -c 1500 gpurun_out/bench_proj_i8_pair.json
FM_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_edge_cases_gpu.py -q -m gpu -k "movie_svd_without_motion or frame_count_boundaries" > gpurun_out/edge_q8_tests.log 2>&1; echo "edge tests (Q6 padding, Q8 movie-only) rc $?"; tail -3 gpurun_out/edge_q8_tests.log
# structured operands that say HOW a kernel is wrong (one process per kernel: a trap poisons the context)
for part in gemm 0 1; do timeout 200 python scripts/triage_i8.py $part > gpurun_out/triage_i8_$part.log 2>&1; echo "triage $part rc $?: $(grep -c "^ok" gpurun_out/triage_i8_$part.log) ok, $(grep -c "^FAIL\|^ERROR" gpurun_out/triage_i8_$part.log) bad"; grep "^FAIL\|^ERROR" gpurun_out/triage_i8_$part.log | head -6; done
timeout 200 python scripts/probe_gemm_i8.py > gpurun_out/probe_gemm_i8.log 2>&1; echo "probe i8 rc $?"; cat gpurun_out/probe_gemm_i8.log | tail -20
for mode in "tf32x3 0" "sliced_merge 0" "sliced 0" "sliced 1" "i8_merge 0" "i8 0"; do
  set -- $mode
  FM_GRAM_MODE=$1 FM_SLICED_SINGLE=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_gram_$1_$2.json 2> gpurun_out/bench_gram_$1_$2.err; echo "bench $1 single=$2 rc $?"
  python - "$1" "$2" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench_gram_{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], "single", sys.argv[2], "ms/step", round(d["ms_per_step"], 2), {k: round(v, 2) for k, v in d["stage_ms"].items() if ".gram" in k})
PY
done
# ncu evidence for the integer projection, only once its tests have passed without a profiler attached
if grep -q " passed" gpurun_out/project_i8_tests.log && ! grep -q "failed" gpurun_out/project_i8_tests.log; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:project_i8_kernel -s 1 -c 1 -f -o gpurun_out/prof_project_i8 python scripts/probe_project_i8.py > gpurun_out/ncu_project_i8.log 2>&1; echo "ncu project_i8 rc $?"
  ncu -i gpurun_out/prof_project_i8.ncu-rep --page raw --csv > gpurun_out/prof_project_i8_raw.csv 2>/dev/null
  grep -E "sm__inst_executed_pipe_tensor|sm__pipe_tensor|dram__bytes_(read|write).sum\"|gpu__time_duration.sum|l1tex__data_bank_conflicts|smsp__warp_issue_stalled" gpurun_out/prof_project_i8_raw.csv | head -20
fi
