#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for combo in "9 9 9" "9 9 0" "9 0 0" "0 9 0" "0 0 0"; do
  set -- $combo
  echo "=== GRAM=$1 LEFT=$2 PROJ=$3"
  FM_GRAM_IMPL=$1 FM_LEFT_IMPL=$2 FM_PROJ_IMPL=$3 timeout 300 python scripts/probe_pipeline.py 0 0 2>&1 | grep -E "S rel err|orth|trace" | cut -c1-230
done
