#!/bin/bash
# gpurun payload: full GPU suite (with the edge-shape cases) + smoke.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread > gpurun_out/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -40 gpurun_out/pytest_gpu.txt
