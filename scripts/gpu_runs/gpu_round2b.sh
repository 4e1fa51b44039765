#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -3 gpurun_out/bench_n1.err
for f in gpurun_out/bench_n1; do grep -o '"value": [0-9.]*' $f.json | head -1; grep -o '"ms_per_step": [0-9.]*' $f.json | head -1; grep -o '"host_ms_per_step[^]]*]' $f.json; grep -o "\"report_ms_each_step\": \[[^a-z]*\]" $f.json; grep -o "\"host_marks_ms_slowest_step\": {[^}]*}" $f.json; grep -o "\"host_marks_ms_fastest_step\": {[^}]*}" $f.json; done
