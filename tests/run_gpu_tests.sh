#!/bin/bash
# Round-trip script for a gpurun call: full GPU test-suite with per-test timeouts; logs land in gpurun_out/.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -q -m gpu --timeout 120 --timeout-method=thread "$@" > gpurun_out/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -60 gpurun_out/pytest_gpu.txt
echo "--- parity report"; cat gpurun_out/parity_report.jsonl 2>/dev/null
