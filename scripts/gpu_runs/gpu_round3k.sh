#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python bench.py --no-cpu > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -2 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_default.json").read().strip().splitlines()[-1])
print("value",d["value"],"ms/step",d["ms_per_step"],"host",d["host_ms_per_step"],"e2e",d["e2e"]["value"] if d["e2e"] else None)
PY
