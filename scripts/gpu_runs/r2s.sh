#!/bin/bash
# round 2, call s: persistent pinned result pool (pipeline tests, bench line), then the ncu launch list of bench.py's timed region
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_run_config.py tests/test_save_parity.py -x -q -m gpu > gpurun_out/r2s_tests.log 2>&1; echo "tests rc $?"; tail -4 gpurun_out/r2s_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2s.out 2> gpurun_out/bench_r2s.err; echo "bench rc $?"; tail -c 400 gpurun_out/bench_r2s.err; tail -n 1 gpurun_out/bench_r2s.out > gpurun_out/bench_r2s.json; cut -c1-400 gpurun_out/bench_r2s.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2s_launches_bench_steps2_warmup3.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/r2s_ncu.log 2>&1; echo "ncu rc $?"; tail -c 300 gpurun_out/r2s_ncu.log; wc -l gpurun_out/r2s_launches_bench_steps2_warmup3.csv
