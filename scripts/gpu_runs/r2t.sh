#!/bin/bash
# round 2, call t: output file primed during the pass (run / save tests, bench line with e2e)
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_run_config.py tests/test_save_parity.py tests/test_reference_suite_gpu.py tests/test_video_files.py -x -q -m gpu > gpurun_out/r2t_tests.log 2>&1; echo "tests rc $?"; tail -4 gpurun_out/r2t_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 --e2e-steps 5 > gpurun_out/bench_r2t.out 2> gpurun_out/bench_r2t.err; echo "bench rc $?"; tail -c 400 gpurun_out/bench_r2t.err; tail -n 1 gpurun_out/bench_r2t.out > gpurun_out/bench_r2t.json; cut -c1-300 gpurun_out/bench_r2t.json
