#!/bin/bash
# Round-2 opener: validate the sliced merge Gram (written without a GPU at the end of round 1) and time it.
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
FM_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_sliced_gram_gpu.py -x -q -m gpu -s > gpurun_out/sliced_tests.log 2>&1; echo "sliced tests rc $?"
tail -15 gpurun_out/sliced_tests.log
FM_GRAM_MODE=sliced timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_sliced.json 2> gpurun_out/bench_sliced.err; echo "bench sliced rc $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_sliced.json").read().strip().splitlines()[-1])
print("sliced merge Gram: ms/step", d["ms_per_step"], {k: v for k, v in d["stage_ms"].items() if ".gram" in k})
PY
