#!/bin/bash
# second GPU contact: full gpu test suite, accuracy probe, eig variants, bench, ncu evidence
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc $?" >> gpurun_out/summary.log
timeout 600 python scripts/probe_pipeline.py 0 0 > gpurun_out/pipe_impl0.log 2>&1; echo "pipe rc $?" >> gpurun_out/summary.log
timeout 600 python scripts/probe_eig.py > gpurun_out/eig.log 2>&1; echo "eig rc $?" >> gpurun_out/summary.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc $?" >> gpurun_out/summary.log
timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc $?" >> gpurun_out/summary.log
timeout 300 python scripts/profile_kernels.py > gpurun_out/profile_plain.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:bin_motion -s 2 -c 1 -f -o gpurun_out/prof_k1 python scripts/profile_kernels.py > gpurun_out/ncu_k1.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 2 -c 1 -f -o gpurun_out/prof_gemm python scripts/profile_kernels.py > gpurun_out/ncu_gemm.log 2>&1
echo "ncu full rc $?" >> gpurun_out/summary.log
timeout 900 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err && \
  timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bin_motion|gemm_|splitk|gram_assemble|row_means|topk|transpose|mean_update|gather_roi|f64_to|f32_to" -s 620 -c 260 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json | cut -c1-1500; tail -3 gpurun_out/bench.err
