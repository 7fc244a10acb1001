// Device-side parameter block of the tcgen05 GEMM kernel (gemm.cu: CTA pairs, cta_group::2, 256-row tiles).
#pragma once
#include <cuda_bf16.h>

namespace etp {

struct GemmDev {
  int M, N, K;
  int tiles_m, tiles_n, k_splits, kb_per_split;  // kb = 64-wide k blocks
  float alpha;
  const float* bias;
  int act;       // 0 none, 1 gelu(erf), 2 relu
  int aux_mode;  // 0 none, 1: *= gelu'(aux), 2: *= (aux > 0), 3: *= aux
  const __nv_bfloat16* aux;
  int ld_aux;
  const float* resid;
  int ld_resid;
  float* out_f32;
  int ld_f32;
  int atomic;
  __nv_bfloat16* out_bf16;
  int ld_bf16;
  __nv_bfloat16* out_pre;  // pre-activation copy (bf16), for the GELU backward
  int ld_pre;
  int pre_mode;   // 0: out_pre = pre-activation; 1: out_pre = gelu'(pre-activation)
  float* colsum;  // fp32 [N] += column sums of the final value
  uint32_t drop_key, drop_thr;  // dropout of the activated value before the residual add (thr 0 = off)
  float drop_scale;
};

}  // namespace etp
