#!/bin/bash
# Round artefacts for profiles/: (1) ncu launch list of the bench command, (2) full capture of the GEMM kernel
# (grouped weight-gradient launch and the FFN GEMMs), (3) N-GPU bench if more than one device is visible.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 700 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 150 -c 12 -o gpurun_out/prof_gemm -f \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.txt 2>&1
tail -2 gpurun_out/ncu_full.txt; wc -l gpurun_out/launches.csv
