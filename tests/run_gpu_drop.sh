#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dropout_gpu.py -q -m gpu -x --timeout 100 --timeout-method=thread -s > gpurun_out/pytest_drop.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_drop.txt
tail -30 gpurun_out/pytest_drop.txt
if grep -q "pytest exit 0" gpurun_out/pytest_drop.txt; then
  bash tests/run_gpu_tests.sh | tail -6
  timeout 600 python bench.py --steps 20 --warmup 5 --kernel-report gpurun_out/kernel_report.txt --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err
  echo "bench exit $?" >> gpurun_out/bench.err
  cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
fi
