#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
N=${1:-2}
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for impl in 7 8; do timeout 180 python scripts/probe_gemm.py $impl > gpurun_out/gemm_impl$impl.log 2>&1; echo "gemm impl $impl rc $?" >> gpurun_out/summary.log; done
timeout 180 python scripts/probe_gemm.py 0 > gpurun_out/gemm_impl0.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc $?" >> gpurun_out/summary.log
FM_BENCH_DIAG=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -4 gpurun_out/pytest_gpu.log | cut -c1-200; tail -14 gpurun_out/gemm_impl7.log; tail -6 gpurun_out/gemm_impl8.log;  tail -5 gpurun_out/gemm_impl0.log; grep DIAG gpurun_out/bench_n$N.err; cut -c1-300 gpurun_out/bench_n$N.json; grep -o '"host_ms_per_step[^]]*]' gpurun_out/bench_n$N.json
