#!/bin/bash
# N GPUs: multi-GPU correctness + bench at N
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
N=${1:-2}
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc $?" >> gpurun_out/summary.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/check_dist.py > gpurun_out/check_dist.log 2>&1; echo "check_dist rc $?" >> gpurun_out/summary.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench n$N rc $?" >> gpurun_out/summary.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -8 gpurun_out/pytest_gpu.log | cut -c1-200; tail -12 gpurun_out/check_dist.log; cut -c1-400 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err; cut -c1-300 gpurun_out/bench_n1.json
