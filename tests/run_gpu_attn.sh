#!/bin/bash
# attention kernels: ops tests (short timeouts: a hung kernel must not eat the call), then full tests + bench
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_ops_gpu.py -q -m gpu -x --timeout 40 --timeout-method=thread -k "attention" > gpurun_out/pytest_ops.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_ops.txt
tail -15 gpurun_out/pytest_ops.txt
if grep -q "pytest exit 0" gpurun_out/pytest_ops.txt; then
  bash tests/run_gpu_tests.sh | tail -8
  timeout 600 python bench.py --steps 20 --warmup 5 --kernel-report gpurun_out/kernel_report.txt --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?" >> gpurun_out/bench.err
  cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; head -30 gpurun_out/kernel_report.txt | cut -c1-40,150-260
fi
