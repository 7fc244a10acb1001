#!/bin/bash
# gpurun payload: new-row tests, secondary bench workloads, the main bench line, ncu launch list + GEMM full capture.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl
timeout 600 python -m pytest tests/test_pretrain_gpu.py tests/test_packing_gpu.py tests/test_shapes_gpu.py -q -m gpu --timeout 300 --timeout-method=thread > gpurun_out/pytest_new.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_new.txt
tail -40 gpurun_out/pytest_new.txt
timeout 400 python bench.py --workload pretrain --steps 10 --warmup 4 > gpurun_out/bench_pretrain.json 2> gpurun_out/bench_pretrain.err
tail -3 gpurun_out/bench_pretrain.err; cat gpurun_out/bench_pretrain.json
timeout 200 python bench.py --workload packing --steps 10 --warmup 3 > gpurun_out/bench_packing.json 2> gpurun_out/bench_packing.err
tail -3 gpurun_out/bench_packing.err; cat gpurun_out/bench_packing.json
timeout 400 python bench.py --steps 20 --warmup 5 --kernel-report gpurun_out/kernel_report.txt > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err
tail -3 gpurun_out/bench_main.err; cat gpurun_out/bench_main.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 700 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.txt 2>&1
timeout 420 ncu --set full --clock-control none -k regex:gemm_tcgen05 -s 330 -c 60 -o gpurun_out/prof_gemm_step -f \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.txt 2>&1
tail -2 gpurun_out/ncu_full.txt; wc -l gpurun_out/launches.csv; ls -la gpurun_out/*.ncu-rep
echo "--- parity report"; cat gpurun_out/parity_report.jsonl 2>/dev/null | tail -12
