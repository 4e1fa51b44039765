#!/bin/bash
# round 2, call z6: research probe — the merge eigenproblem by block subspace iteration (pieces and whole loop on B200)
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 280 python scripts/probe_merge_subspace.py > gpurun_out/probe_merge_subspace.log 2>&1; echo "probe rc $?"; tail -12 gpurun_out/probe_merge_subspace.log | cut -c1-700
