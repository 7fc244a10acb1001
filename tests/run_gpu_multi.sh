#!/bin/bash
# gpurun --gpus N payload: data-parallel bench at N ranks (torchrun, NCCL), both arms.
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "exit $?" >> gpurun_out/bench_n$N.err
cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
