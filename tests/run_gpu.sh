#!/bin/bash
# gpurun payload (one script, parametrised):   bash tests/run_gpu.sh <what> [args...]
#   tests [pytest args]   pytest -m gpu (default: whole suite) -> gpurun_out/pytest_gpu.txt
#   bench [bench args]    python bench.py ... -> gpurun_out/bench.json (+ .err)
#   report [bench args]   bench with --kernel-report -> gpurun_out/kernel_report.txt
#   launches              ncu launch list of the bench command -> gpurun_out/launches.csv
#   ncu <kernel-regex> [skip] [count]   ncu --set full capture -> gpurun_out/prof_<regex>.ncu-rep
#   metrics               ncu per-launch DRAM bytes / tensor-pipe / duration over one train step -> gpurun_out/step_metrics.csv
#   smoke                 __graft_entry__.smoke()
#   benchn <N> <name> [bench args]   torchrun over N GPUs of one box -> gpurun_out/bench_<name>.json
#   peer <N>              fused peer-memory update vs the NCCL path on N GPUs (tests/dp_peer_worker.py) -> gpurun_out/peer_n<N>.json
# Several <what> can be chained with '+' between argument groups:  bash tests/run_gpu.sh tests -k precision + bench --steps 20
mkdir -p gpurun_out
run_one() {
  what=$1; shift
  case "$what" in
    tests)
      if [ $# -eq 0 ]; then set -- tests; fi
      timeout 2400 python -m pytest -m gpu -q "$@" > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.txt ;;
    bench)
      timeout 1500 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err ;;
    benchto)   # benchto <name> [bench args]: quick line (no CPU / eager / soak legs) -> gpurun_out/bench_<name>.json
      name=$1; shift
      timeout 900 python bench.py --no-cpu-baseline --no-gpu-eager --no-soak "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench $name rc=$?"; cat gpurun_out/bench_$name.json; tail -2 gpurun_out/bench_$name.err ;;
    benchn)    # benchn <N> <name> [bench args]: torchrun over N GPUs of this box
      n=$1; name=$2; shift; shift
      timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n --no-cpu-baseline --no-gpu-eager "$@" > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err; echo "bench $name rc=$?"; cat gpurun_out/bench_$name.json; tail -2 gpurun_out/bench_$name.err ;;
    peer)      # peer <N>: the fused peer-memory update against the NCCL path on N GPUs (tests/dp_peer_worker.py)
      n=$1; shift
      timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29561 tests/dp_peer_worker.py > gpurun_out/peer_n$n.json 2> gpurun_out/peer_n$n.err; echo "peer n=$n rc=$?"; cat gpurun_out/peer_n$n.json; tail -5 gpurun_out/peer_n$n.err ;;
    report)
      timeout 900 python bench.py --no-cpu-baseline --no-gpu-eager --kernel-report gpurun_out/kernel_report.txt "$@" > gpurun_out/bench_report.json 2> gpurun_out/bench_report.err
      echo "report rc=$?"; cat gpurun_out/bench_report.json; head -40 gpurun_out/kernel_report.txt ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 700 --csv --log-file gpurun_out/launches.csv \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-soak "$@" > gpurun_out/bench_under_ncu.txt 2>&1; wc -l gpurun_out/launches.csv ;;
    ncu)
      k=$1; skip=${2:-100}; cnt=${3:-8}
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c $cnt -o gpurun_out/prof_$k -f \
        python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-eager --no-soak > gpurun_out/ncu_full.txt 2>&1; tail -2 gpurun_out/ncu_full.txt ;;
    metrics)
      timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed \
        --clock-control none -s 2000 -c 260 --csv --log-file gpurun_out/step_metrics.csv \
        python bench.py --steps 1 --warmup 8 --no-cpu-baseline --no-gpu-eager --no-soak > gpurun_out/bench_under_ncu2.txt 2>&1; wc -l gpurun_out/step_metrics.csv ;;
    smoke)
      timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.txt ;;
    *) echo "unknown: $what"; return 2 ;;
  esac
}
args=()
for a in "$@"; do
  if [ "$a" = "+" ]; then run_one "${args[@]}"; args=(); else args+=("$a"); fi
done
[ ${#args[@]} -gt 0 ] && run_one "${args[@]}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > gpurun_out/smi_end.txt 2>&1
exit 0
