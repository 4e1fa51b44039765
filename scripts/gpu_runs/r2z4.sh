#!/bin/bash
# round 2, call z4: the whole GPU suite and smoke() on the round's last code
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2z4_smoke.log 2>&1; echo "smoke rc $?"; tail -1 gpurun_out/r2z4_smoke.log
timeout 1800 python -m pytest tests -q -m gpu --durations=5 > gpurun_out/gputests_r2z4.log 2>&1; echo "gpu tests rc $?"; tail -9 gpurun_out/gputests_r2z4.log | cut -c1-250
