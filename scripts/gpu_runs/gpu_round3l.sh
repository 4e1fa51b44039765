#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
timeout 80 python scripts/bench_ingest.py --frames 10000 --cache-only > gpurun_out/ingest_default_workers.json 2> gpurun_out/ingest2.err; echo "ingest rc $?"
cat gpurun_out/ingest_default_workers.json; tail -3 gpurun_out/ingest2.err
