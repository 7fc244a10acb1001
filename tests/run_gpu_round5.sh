#!/bin/bash
# gpurun payload: full GPU suite on the tree with the leaner GEMM epilogue, main bench + kernel report, bench lines at the
# other BASELINE.json shapes.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread > gpurun_out/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -6 gpurun_out/pytest_gpu.txt
timeout 400 python bench.py --steps 20 --warmup 5 --kernel-report gpurun_out/kernel_report.txt > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err
tail -3 gpurun_out/bench_main.err; cat gpurun_out/bench_main.json
bash tests/run_gpu_round4.sh
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
