#!/bin/bash
# gpurun payload: full GPU suite on the final tree, packing bench, ncu evidence (launch list, per-kernel metrics of one
# whole step, one --set full capture of a few GEMM launches exported as CSV), main bench with the kernel report.
# Large ncu reports stay in /tmp on the box: gpurun_out/ must stay under 64 MiB.
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.jsonl gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests -q -m gpu --timeout 300 --timeout-method=thread > gpurun_out/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.txt
tail -15 gpurun_out/pytest_gpu.txt
timeout 200 python bench.py --workload packing --steps 10 --warmup 3 > gpurun_out/bench_packing.json 2> gpurun_out/bench_packing.err
tail -3 gpurun_out/bench_packing.err; cat gpurun_out/bench_packing.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 700 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.txt 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed
timeout 400 ncu --metrics $M --clock-control none -s 700 -c 235 --csv --log-file gpurun_out/step_metrics.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu2.txt 2>&1
timeout 400 ncu --set full --clock-control none -k regex:gemm_tcgen05 -s 330 -c 8 -o /tmp/prof_gemm8 -f \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.txt 2>&1
ncu -i /tmp/prof_gemm8.ncu-rep --page raw --csv > gpurun_out/gemm_full_raw.csv 2>/dev/null
ncu -i /tmp/prof_gemm8.ncu-rep --page details --csv > gpurun_out/gemm_full_details.csv 2>/dev/null
tail -2 gpurun_out/ncu_full.txt; wc -l gpurun_out/launches.csv gpurun_out/step_metrics.csv gpurun_out/gemm_full_raw.csv
timeout 400 python bench.py --steps 20 --warmup 5 --kernel-report gpurun_out/kernel_report.txt > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err
tail -3 gpurun_out/bench_main.err; cat gpurun_out/bench_main.json
du -sh gpurun_out
