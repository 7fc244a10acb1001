#!/bin/bash
# same-box A/B of the GEMM epilogue-warp policy: ETP_GEMM_EW = 8 | 16 | 1 (16 only with dropout) | unset (default policy)
mkdir -p gpurun_out
for ew in 8 16 1 0; do
  export ETP_GEMM_EW=$ew
  [ "$ew" = "0" ] && unset ETP_GEMM_EW
  echo "=== EW policy $ew"
  timeout 120 python tests/diag_gemm.py epi 2>&1 | grep "^epi" | awk '{print $(NF-3)}' | tr '\n' ' '; echo
  for dr in on off; do
    timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --dropout $dr 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dropout $dr:', round(d['value'],1), round(d['ms_per_step'],3), 'gemm', round(d['roofline']['gemm_ms_per_step'],3), round(d['roofline']['achieved'],0))"
  done
done
