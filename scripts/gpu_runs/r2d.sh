#!/bin/bash
# round 2, call d: first run of the hand-written batched eigensolver (csrc/eigh_cluster.cu)
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 300 python -m pytest tests/test_eigh_cluster_gpu.py -q -m gpu -x > gpurun_out/eigh_tests.log 2>&1; echo "eigh tests rc $?"; tail -30 gpurun_out/eigh_tests.log | cut -c1-300
timeout 300 python scripts/probe_eigh_cluster.py > gpurun_out/probe_eigh_cluster.log 2>&1; echo "probe rc $?"; tail -12 gpurun_out/probe_eigh_cluster.log | cut -c1-300
timeout 300 python -m pytest tests/test_video_files.py tests/test_pipeline_gpu.py -q -m gpu > gpurun_out/r2d_tests.log 2>&1; echo "video + pipeline tests rc $?"; tail -8 gpurun_out/r2d_tests.log | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-extras > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo "bench rc $?"; tail -c 2500 gpurun_out/bench_r2d.json
