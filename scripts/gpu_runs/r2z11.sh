#!/bin/bash
# round 2, call z11: smoke() on the round's last commit
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 60 python __graft_entry__.py smoke > gpurun_out/r2z11_smoke.log 2>&1; echo "smoke rc $?"; tail -2 gpurun_out/r2z11_smoke.log
