#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 55 python -m pytest tests/test_k1_gpu.py tests/test_reference_suite_gpu.py -x -q -m gpu -k "motion_ref or test_output" > gpurun_out/pytest_last.log 2>&1; echo "rc $?"
tail -5 gpurun_out/pytest_last.log
