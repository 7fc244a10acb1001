#!/bin/bash
# GEMM bring-up script: every stage under its own timeout so a hung kernel cannot eat the call.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
rm -f gpurun_out/diag_gemm.txt
for g in basic kk bmn amn; do
  echo "=== pair group $g" >> gpurun_out/diag_gemm.txt
  timeout 120 python tests/diag_gemm.py $g >> gpurun_out/diag_gemm.txt 2>&1
  echo "exit $?" >> gpurun_out/diag_gemm.txt
done
echo "=== pair perf" >> gpurun_out/diag_gemm.txt
timeout 200 python tests/diag_gemm.py perf >> gpurun_out/diag_gemm.txt 2>&1
echo "exit $?" >> gpurun_out/diag_gemm.txt
timeout 200 python tests/diag_gemm.py epi >> gpurun_out/diag_gemm.txt 2>&1
timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -x --timeout 60 --timeout-method=thread > gpurun_out/pytest_ops.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_ops.txt
grep -v "^  \|^   " gpurun_out/diag_gemm.txt | tail -80
tail -15 gpurun_out/pytest_ops.txt
if grep -q "pytest exit 0" gpurun_out/pytest_ops.txt; then
  bash tests/run_gpu_tests.sh
  timeout 600 python bench.py --steps 20 --warmup 5 --kernel-report gpurun_out/kernel_report.txt > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?" >> gpurun_out/bench.err
  cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; head -40 gpurun_out/kernel_report.txt
fi
