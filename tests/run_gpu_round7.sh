#!/bin/bash
# gpurun payload: full GPU suite + main bench (no CPU leg) with the kernel report: A/B of the LayerNorm-backward rewrite.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 600 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread > gpurun_out/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -25 gpurun_out/pytest_gpu.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --kernel-report gpurun_out/kernel_report.txt > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err
tail -3 gpurun_out/bench_main.err; cat gpurun_out/bench_main.json
grep -i "layernorm" gpurun_out/kernel_report.txt | cut -c1-60
