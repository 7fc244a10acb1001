#!/bin/bash
# gpurun payload: smoke, bench (both arms), ncu launch list of the bench command, optional full capture of the GEMM.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.txt
timeout 600 python bench.py --steps 20 --warmup 5 "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_under_ncu.txt 2>&1
if [ -n "$NCU_FULL" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 40 -c 3 -o gpurun_out/prof_gemm \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/ncu_full.txt 2>&1
fi
tail -3 gpurun_out/smoke.txt; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err; wc -l gpurun_out/launches.csv
