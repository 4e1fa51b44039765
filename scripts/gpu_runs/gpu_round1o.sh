#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc $?" >> gpurun_out/summary.log
timeout 900 python bench.py --steps 1 --warmup 5 --no-e2e --no-cpu > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err && \
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bin_motion|gemm_|splitk|gram_assemble|row_means|row_dot|topk|transpose_kernel|mean_update|gather_roi|split_lo|sign_from" -s 1500 -c 330 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 5 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300; tail -3 gpurun_out/ncu_launches.log | cut -c1-300; wc -l gpurun_out/launches.csv; grep -o '"gpu_launches": [0-9]*' gpurun_out/bench_short.json
