// Device-side parameter block shared by the tcgen05 GEMM kernels (gemm.cu: one CTA per 128-row tile;
// gemm_pair.cu: CTA pairs, cta_group::2, 256-row tiles).
#pragma once
#include <cuda_bf16.h>

namespace etp {

struct GemmDev {
  int M, N, K;
  int tiles_m, tiles_n, k_splits, kb_per_split;  // kb = 64-wide k blocks
  float alpha;
  const float* bias;
  int act;       // 0 none, 1 gelu(erf), 2 relu
  int aux_mode;  // 0 none, 1: *= gelu'(aux), 2: *= (aux > 0)
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
  float* colsum;  // fp32 [N] += column sums of the final value (CTA-pair kernel only)
};

struct GemmArgs;
// CTA-pair kernel (gemm_pair.cu).  `d` arrives with the epilogue fields filled; tiling fields are set here.
int gemm_pair(const GemmArgs& a, GemmDev d, cudaStream_t stream);

}  // namespace etp
