#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_fullres_gpu.py tests/test_k1_gpu.py -q -m gpu > gpurun_out/pytest_fullres.log 2>&1; echo "pytest fullres+k1 rc $?" >> gpurun_out/summary.log
timeout 900 python -m pytest tests -x -q -m gpu --deselect tests/test_fullres_gpu.py --deselect tests/test_k1_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest gpu rc $?" >> gpurun_out/summary.log
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu --no-e2e --with-rois > gpurun_out/bench_rois.json 2> gpurun_out/bench_rois.err; echo "bench rois rc $?" >> gpurun_out/summary.log
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench n1 rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -5 gpurun_out/pytest_fullres.log; tail -3 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench_rois.err
python - <<'PY'
import json
for f in ("bench_rois","bench_n1"):
    d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f,"value",d["value"],"ms/step",d["ms_per_step"],"host",d["host_ms_per_step"], "k1 roofline", d["roofline"]["frac"])
    print({k:v for k,v in d["stage_ms"].items() if "stage3" in k or "fullres" in k})
PY
