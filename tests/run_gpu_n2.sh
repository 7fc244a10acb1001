#!/bin/bash
# gpurun --gpus 2 payload: data-parallel bench at 2 ranks (torchrun, NCCL), B200 arm only, short.
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "exit $?" >> gpurun_out/bench_n2.err
cat gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
