#!/bin/bash
# gpurun payload: the pre-training twin's GPU tests first (no -x: every failure is reported), then the rest of the suite.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 600 python -m pytest tests/test_pretrain_gpu.py -q -m gpu --timeout 300 --timeout-method=thread > gpurun_out/pytest_pretrain.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_pretrain.txt
tail -80 gpurun_out/pytest_pretrain.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 120 --timeout-method=thread --deselect tests/test_pretrain_gpu.py > gpurun_out/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -15 gpurun_out/pytest_gpu.txt
echo "--- parity report"; cat gpurun_out/parity_report.jsonl 2>/dev/null | tail -20
