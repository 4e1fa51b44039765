#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for kc in 128 192 320; do
  echo "=== KCHAIN_GRAM=$kc (impl 0 everywhere)"
  FM_KCHAIN_GRAM=$kc timeout 300 python scripts/probe_pipeline.py 0 0 2>&1 | grep -E "S rel err|orth|device ms" | cut -c1-330
done
FM_KCHAIN_GRAM=128 timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"; grep "^FAILED" gpurun_out/pytest_gpu.log | cut -c1-160; tail -2 gpurun_out/pytest_gpu.log
