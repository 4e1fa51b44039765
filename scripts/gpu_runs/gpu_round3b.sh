#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_fullres_gpu.py -q -m gpu > gpurun_out/pytest_fullres.log 2>&1; echo "pytest fullres rc $?" >> gpurun_out/summary.log
timeout 300 python scripts/bench_fullres.py > gpurun_out/fullres_bench.json 2> gpurun_out/fullres_bench.err; echo "bench fullres rc $?" >> gpurun_out/summary.log
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu --no-e2e --with-rois > gpurun_out/bench_rois.json 2> gpurun_out/bench_rois.err; echo "bench rois rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -5 gpurun_out/pytest_fullres.log; cat gpurun_out/fullres_bench.json; tail -3 gpurun_out/fullres_bench.err; tail -3 gpurun_out/bench_rois.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_rois.json").read().strip().splitlines()[-1])
print("value",d["value"],"ms/step",d["ms_per_step"],"host",d["host_ms_per_step"])
print({k:v for k,v in d["stage_ms"].items() if "stage3" in k or "fullres" in k})
PY
