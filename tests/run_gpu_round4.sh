#!/bin/bash
# gpurun payload: bench lines at the other BASELINE.json shapes (per-GPU slices of c4 / c5, and c2 forward), N = 1.
mkdir -p gpurun_out
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch 64 --nodes 80 --tokens 80 --x-layers 4 > gpurun_out/bench_c4_n1.json 2> gpurun_out/bench_c4.err
tail -2 gpurun_out/bench_c4.err; cat gpurun_out/bench_c4_n1.json
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --batch 32 --nodes 120 --tokens 512 --x-layers 4 > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5.err
tail -2 gpurun_out/bench_c5.err; cat gpurun_out/bench_c5_n1.json
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --mode fwd --batch 32 --nodes 40 --tokens 160 --x-layers 4 > gpurun_out/bench_c2_fwd.json 2> gpurun_out/bench_c2.err
tail -2 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2_fwd.json
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --mode fwd > gpurun_out/bench_c3_fwd.json 2> gpurun_out/bench_c3f.err
tail -2 gpurun_out/bench_c3f.err; cat gpurun_out/bench_c3_fwd.json
