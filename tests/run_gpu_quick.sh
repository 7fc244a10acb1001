#!/bin/bash
# quick GEMM A/B: perf + epilogue cases for the pair kernel, then ops tests + full GPU tests + bench with kernel report
mkdir -p gpurun_out
rm -f gpurun_out/diag_gemm.txt
echo "=== pair perf" >> gpurun_out/diag_gemm.txt
timeout 200 python tests/diag_gemm.py perf >> gpurun_out/diag_gemm.txt 2>&1
timeout 200 python tests/diag_gemm.py epi >> gpurun_out/diag_gemm.txt 2>&1
grep -v "^  8x8\|^   " gpurun_out/diag_gemm.txt | grep -v cuBLAS | tail -40
timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -x --timeout 60 --timeout-method=thread > gpurun_out/pytest_ops.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_ops.txt
tail -5 gpurun_out/pytest_ops.txt
if grep -q "pytest exit 0" gpurun_out/pytest_ops.txt; then
  bash tests/run_gpu_tests.sh | tail -8
  timeout 600 python bench.py --steps 20 --warmup 5 --kernel-report gpurun_out/kernel_report.txt --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?" >> gpurun_out/bench.err
  cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; head -45 gpurun_out/kernel_report.txt | cut -c1-40,150-260
fi
