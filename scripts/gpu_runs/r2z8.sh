#!/bin/bash
# round 2, call z8: motion + movie SVD at 3 750 stacked components, merge solve on cuSOLVER against subspace iteration
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 170 python scripts/probe_merge_modes.py > gpurun_out/probe_merge_modes.log 2>&1; echo "probe rc $?"; tail -22 gpurun_out/probe_merge_modes.log | cut -c1-300
