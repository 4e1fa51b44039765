#!/bin/bash
# first GPU contact: environment, K1 parity, GEMM bring-up, pipeline with both GEMM paths
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
{ nvidia-smi; nproc; free -g; python -c "import os;print('cpus',os.cpu_count())"; } > gpurun_out/env.log 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_k1_gpu.py -q -x > gpurun_out/k1.log 2>&1; echo "k1 rc $?" >> gpurun_out/summary.log
for impl in 1 0 2 3 4; do
  timeout 300 python scripts/probe_gemm.py $impl > gpurun_out/gemm_impl$impl.log 2>&1; echo "gemm impl $impl rc $?" >> gpurun_out/summary.log
done
timeout 600 python scripts/probe_eig.py > gpurun_out/eig.log 2>&1; echo "eig rc $?" >> gpurun_out/summary.log
timeout 900 python scripts/probe_pipeline.py 1 0 > gpurun_out/pipe_impl1.log 2>&1; echo "pipe impl1 rc $?" >> gpurun_out/summary.log
timeout 900 python scripts/probe_pipeline.py 0 0 > gpurun_out/pipe_impl0.log 2>&1; echo "pipe impl0 rc $?" >> gpurun_out/summary.log
timeout 900 python scripts/probe_pipeline.py 0 1 > gpurun_out/pipe_impl0_eig32.log 2>&1; echo "pipe impl0 eig32 rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log
tail -5 gpurun_out/k1.log; tail -12 gpurun_out/gemm_impl0.log; tail -20 gpurun_out/pipe_impl1.log
