#!/bin/bash
# ncu captures (full set, source-level) of the kernels named in $1 (regex), $2 launches after skipping $3; bench payload
mkdir -p gpurun_out
K=${1:-attention_bwd_tc}; C=${2:-2}; S=${3:-20}; NAME=${4:-prof}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C -o gpurun_out/$NAME -f \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$NAME.txt 2>&1
echo "ncu exit $?" >> gpurun_out/ncu_$NAME.txt
tail -3 gpurun_out/ncu_$NAME.txt; ls -la gpurun_out/$NAME.ncu-rep
