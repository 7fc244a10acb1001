#!/bin/bash
# gpurun payload: full GPU test-suite, then smoke + bench + ncu launch list.
bash tests/run_gpu_tests.sh "$@"
bash tests/run_gpu_bench.sh
