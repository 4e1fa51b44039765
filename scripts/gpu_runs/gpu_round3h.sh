#!/bin/bash
mkdir -p gpurun_out; cd "$(dirname "$0")/../.."
rm -f gpurun_out/summary.log
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_video_files.py tests/test_reference_suite_gpu.py -q -m gpu > gpurun_out/pytest_video.log 2>&1; echo "pytest video rc $?" >> gpurun_out/summary.log
timeout 600 python scripts/bench_ingest.py --frames 10000 > gpurun_out/ingest.json 2> gpurun_out/ingest.err; echo "ingest rc $?" >> gpurun_out/summary.log
cat gpurun_out/summary.log; tail -15 gpurun_out/pytest_video.log; cat gpurun_out/ingest.json; tail -5 gpurun_out/ingest.err
