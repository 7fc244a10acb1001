#!/bin/bash
# gpurun payload: full GPU test-suite, smoke, bench (with the eager-GPU baseline and the per-kernel table)
mkdir -p gpurun_out
bash tests/run_gpu_tests.sh | tail -6
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.txt
tail -2 gpurun_out/smoke.txt
timeout 900 python bench.py --steps 20 --warmup 5 --gpu-eager --kernel-report gpurun_out/kernel_report.txt "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
