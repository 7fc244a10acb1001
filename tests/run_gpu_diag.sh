#!/bin/bash
# First-contact GPU script: every stage under its own timeout so a hung kernel cannot eat the call.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for g in basic kk bmn amn perf; do
  echo "=== group $g" >> gpurun_out/diag_gemm.txt
  timeout 150 python tests/diag_gemm.py $g >> gpurun_out/diag_gemm.txt 2>&1
  echo "exit $?" >> gpurun_out/diag_gemm.txt
done
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x --timeout 120 > gpurun_out/pytest_ops.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_ops.txt
tail -50 gpurun_out/diag_gemm.txt
tail -30 gpurun_out/pytest_ops.txt
