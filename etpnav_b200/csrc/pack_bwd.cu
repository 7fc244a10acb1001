// Backward of the token-packing kernels, the SAP-head tail and the text embedding (pack.cu), plus AdamW.
// Same mapping as the forward: one warp per 768-wide row, 24 values per lane.  Parameter gradients are
// accumulated per CTA in shared memory (shared atomics) and flushed with one global atomicAdd per
// element per CTA; rows are visited in a grid-stride loop so the number of flushes stays ~2 per SM.
#include "common.cuh"
#include "host.h"
#include "ops.h"

namespace etp {

namespace {
constexpr int kH = 768;

ETP_DEVICE int col_of(int i, int lane) { return ((i >> 2) * 32 + lane) * 4 + (i & 3); }

ETP_DEVICE void ld24(const float* p, int lane, float (&v)[24]) {
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(p + (i * 32 + lane) * 4);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
ETP_DEVICE void st24_bf16(bf16* p, int lane, const float (&v)[24]) {
#pragma unroll
  for (int i = 0; i < 6; ++i)
    *reinterpret_cast<uint2*>(p + (i * 32 + lane) * 4) =
        make_uint2(pack_bf16x2(v[4 * i], v[4 * i + 1]), pack_bf16x2(v[4 * i + 2], v[4 * i + 3]));
}
// shared-memory accumulation of a per-lane 24-vector into s[768]
ETP_DEVICE void sacc(float* s, int lane, const float (&v)[24]) {
#pragma unroll
  for (int i = 0; i < 24; ++i) atomicAdd(&s[col_of(i, lane)], v[i]);
}
// LayerNorm backward of one row held across the warp.  In: dy, x (overwritten by xhat), mean, rstd, gamma.
// Out: dx (in dy).  dgam/dbet receive this row's contributions (dy * xhat, dy).
ETP_DEVICE void ln_bwd_row(float (&dy)[24], float (&x)[24], float mean, float rstd, const float (&g)[24],
                           float (&dgam)[24], float (&dbet)[24]) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    x[i] = (x[i] - mean) * rstd;
    dgam[i] = dy[i] * x[i];
    dbet[i] = dy[i];
    dy[i] *= g[i];
    s1 += dy[i];
    s2 += dy[i] * x[i];
  }
  s1 = warp_sum(s1) * (1.0f / kH);
  s2 = warp_sum(s2) * (1.0f / kH);
#pragma unroll
  for (int i = 0; i < 24; ++i) dy[i] = rstd * (dy[i] - s1 - x[i] * s2);
}
// CTA partial -> global gradient.  128-bit vector reductions (4x fewer L2 atomic operations) when the destination
// is 16-byte aligned, which every parameter gradient in the flat buffer is; all-zero quads are skipped.
ETP_DEVICE void flush(float* dst, const float* s, int n) {
  if (dst == nullptr) return;
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (n & 3) == 0) {
    for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4) {
      const float4 v = *reinterpret_cast<const float4*>(s + i);
      if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                     : "memory");
    }
    return;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = s[i];
    if (v != 0.f) atomicAdd(dst + i, v);
  }
}
ETP_DEVICE void zero_smem(float* s, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = 0.f;
}
int grid_for(int rows) {
  int g = (rows + 7) / 8;
  const int cap = 2 * num_sms();
  return g > cap ? cap : (g < 1 ? 1 : g);
}
}  // namespace

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sap_tail_bwd_kernel(const float* __restrict__ dlogits, const float* __restrict__ relu_out,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ w4, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, const uint8_t* __restrict__ visited,
                                                            const uint8_t* __restrict__ valid, int rows, bf16* __restrict__ dpre,
                                                            float* dgamma, float* dbeta, float* dw4, float* db4) {
  griddep_launch();  // PDL (common.cuh): let the next kernel get scheduled ...
  griddep_wait();    // ... and wait for the previous one before touching memory
  extern __shared__ float sm[];  // [3][768] + [1]
  float* s_g = sm; float* s_b = sm + kH; float* s_w = sm + 2 * kH; float* s_b4 = sm + 3 * kH;
  zero_smem(sm, 3 * kH + 1);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float g[24], bt[24], w[24];
  ld24(gamma, lane, g);
  ld24(beta, lane, bt);
  ld24(w4, lane, w);
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const bool dead = (visited && visited[row]) || (valid && !valid[row]);
    const float dl = dead ? 0.f : dlogits[row];
    float r[24], dh[24], x[24], dgm[24], dbt[24], dw[24];
    ld24(relu_out + static_cast<size_t>(row) * kH, lane, r);
    const float mu = mean[row], rs = rstd[row];
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      x[i] = r[i];
      dh[i] = dl * w[i];
      dw[i] = dl * ((r[i] - mu) * rs * g[i] + bt[i]);
    }
    ln_bwd_row(dh, x, mu, rs, g, dgm, dbt);
#pragma unroll
    for (int i = 0; i < 24; ++i) dh[i] = r[i] > 0.f ? dh[i] : 0.f;
    st24_bf16(dpre + static_cast<size_t>(row) * kH, lane, dh);
    if (dl != 0.f) {
      sacc(s_g, lane, dgm);
      sacc(s_b, lane, dbt);
      sacc(s_w, lane, dw);
      if (lane == 0) atomicAdd(s_b4, dl);
    }
  }
  __syncthreads();
  flush(dgamma, s_g, kH);
  flush(dbeta, s_b, kH);
  flush(dw4, s_w, kH);
  flush(db4, s_b4, 1);
}

int sap_tail_bwd(const float* dlogits, const float* relu_out, const float* gamma, const float* beta, const float* w4,
                 const float* mean, const float* rstd, const uint8_t* visited, const uint8_t* valid, int rows,
                 bf16* dpre_bf16, float* dgamma, float* dbeta, float* dw4, float* db4, cudaStream_t stream) {
  if (rows <= 0) return ETP_OK;
  ETP_CHECK_CUDA(launch_pdl(sap_tail_bwd_kernel, dim3(grid_for(rows)), dim3(256), (3 * kH + 1) * sizeof(float), stream, 
      dlogits, relu_out, gamma, beta, w4, mean, rstd, visited, valid, rows, dpre_bf16, dgamma, dbeta, dw4, db4));
  ETP_LAUNCHED();
  return ETP_OK;
}

// ---------------------------------------------------------------------------------------------------
constexpr int kStepCache = 16;  // step ids below this accumulate in shared memory first

__global__ void __launch_bounds__(256) node_pack_bwd_kernel(const float* __restrict__ dx, const int64_t* __restrict__ step_ids,
                                                             const float* __restrict__ pos_fts, const float* __restrict__ pos_lin,
                                                             const float* __restrict__ stats, const float* __restrict__ pos_g,
                                                             int rows, float* dstep_emb, float* dpos_w, float* dpos_b,
                                                             float* dpos_g, float* dpos_bb) {
  griddep_launch();  // PDL (common.cuh): let the next kernel get scheduled ...
  griddep_wait();    // ... and wait for the previous one before touching memory
  extern __shared__ float sm[];
  float* s_w = sm;                       // [768*7]
  float* s_b = s_w + kH * 7;             // [768]
  float* s_g = s_b + kH;                 // [768]
  float* s_bb = s_g + kH;                // [768]
  float* s_step = s_bb + kH;             // [kStepCache][768]
  zero_smem(sm, kH * 10 + kStepCache * kH);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float g[24];
  ld24(pos_g, lane, g);
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    float d[24], x[24], dgm[24], dbt[24];
    ld24(dx + static_cast<size_t>(row) * kH, lane, d);
    const int64_t id = step_ids[row];
    if (dstep_emb) {
      if (id < kStepCache) {
        sacc(s_step + id * kH, lane, d);
      } else {
#pragma unroll
        for (int i = 0; i < 24; ++i) atomicAdd(dstep_emb + id * kH + col_of(i, lane), d[i]);
      }
    }
    ld24(pos_lin + static_cast<size_t>(row) * kH, lane, x);
    ln_bwd_row(d, x, stats[static_cast<size_t>(row) * 2], stats[static_cast<size_t>(row) * 2 + 1], g, dgm, dbt);
    sacc(s_g, lane, dgm);
    sacc(s_bb, lane, dbt);
    sacc(s_b, lane, d);
    float f[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) f[j] = pos_fts[static_cast<size_t>(row) * 7 + j];
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      const int c = col_of(i, lane);
#pragma unroll
      for (int j = 0; j < 7; ++j)
        if (f[j] != 0.f) atomicAdd(&s_w[c * 7 + j], d[i] * f[j]);
    }
  }
  __syncthreads();
  flush(dpos_w, s_w, kH * 7);
  flush(dpos_b, s_b, kH);
  flush(dpos_g, s_g, kH);
  flush(dpos_bb, s_bb, kH);
  flush(dstep_emb, s_step, kStepCache * kH);
}

int node_pack_bwd(const float* dx, const int64_t* step_ids, const float* pos_fts, const float* pos_lin,
                  const float* stats, const float* pos_g, int rows, float* dstep_emb, float* dpos_w, float* dpos_b,
                  float* dpos_g, float* dpos_bb, cudaStream_t stream) {
  if (rows <= 0) return ETP_OK;
  const size_t smem = (kH * 10 + kStepCache * kH) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    ETP_CHECK_CUDA(cudaFuncSetAttribute(node_pack_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  int grid = grid_for(rows);
  if (grid > num_sms()) grid = num_sms();
  ETP_CHECK_CUDA(launch_pdl(node_pack_bwd_kernel, dim3(grid), dim3(256), smem, stream, dx, step_ids, pos_fts, pos_lin, stats, pos_g, rows, dstep_emb, dpos_w,
                                                    dpos_b, dpos_g, dpos_bb));
  ETP_LAUNCHED();
  return ETP_OK;
}

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pano_pack_bwd_kernel(const PanoPackBwdArgs a) {
  griddep_launch();  // PDL (common.cuh): let the next kernel get scheduled ...
  griddep_wait();    // ... and wait for the previous one before touching memory
  extern __shared__ float sm[];
  // [0] out_g [1] out_b [2] img_g [3] img_b [4] dep_g [5] dep_b [6] loc_g [7] loc_b [8] loc_bias [9] tok [10,11] nav [12..15] loc_w
  zero_smem(sm, 16 * kH);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float og[24];
  ld24(a.out_g, lane, og);
  for (int row = blockIdx.x * 8 + warp; row < a.rows; row += gridDim.x * 8) {
    const float* st = a.stats + static_cast<size_t>(row) * 8;
    float d[24], x[24], dgm[24], dbt[24], g[24];
    ld24(a.dx + static_cast<size_t>(row) * kH, lane, d);
    ld24(a.sum_pre + static_cast<size_t>(row) * kH, lane, x);
    ln_bwd_row(d, x, st[6], st[7], og, dgm, dbt);   // d = dsum
    sacc(sm + 0 * kH, lane, dgm);
    sacc(sm + 1 * kH, lane, dbt);
    sacc(sm + 9 * kH, lane, d);
    sacc(sm + (10 + static_cast<int>(a.nav_types[row])) * kH, lane, d);
    float t[24];
    // img branch
#pragma unroll
    for (int i = 0; i < 24; ++i) t[i] = d[i];
    ld24(a.rgb_lin + static_cast<size_t>(row) * kH, lane, x);
    ld24(a.img_g, lane, g);
    ln_bwd_row(t, x, st[0], st[1], g, dgm, dbt);
    sacc(sm + 2 * kH, lane, dgm);
    sacc(sm + 3 * kH, lane, dbt);
    st24_bf16(a.drgb_lin + static_cast<size_t>(row) * kH, lane, t);
    if (a.dep_lin) {
#pragma unroll
      for (int i = 0; i < 24; ++i) t[i] = d[i];
      ld24(a.dep_lin + static_cast<size_t>(row) * kH, lane, x);
      ld24(a.dep_g, lane, g);
      ln_bwd_row(t, x, st[2], st[3], g, dgm, dbt);
      sacc(sm + 4 * kH, lane, dgm);
      sacc(sm + 5 * kH, lane, dbt);
      st24_bf16(a.ddep_lin + static_cast<size_t>(row) * kH, lane, t);
    }
    ld24(a.loc_lin + static_cast<size_t>(row) * kH, lane, x);
    ld24(a.loc_g, lane, g);
    ln_bwd_row(d, x, st[4], st[5], g, dgm, dbt);   // d = dloc_lin
    sacc(sm + 6 * kH, lane, dgm);
    sacc(sm + 7 * kH, lane, dbt);
    sacc(sm + 8 * kH, lane, d);
    const float4 f = *reinterpret_cast<const float4*>(a.loc_fts + static_cast<size_t>(row) * 4);
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      float* w = sm + 12 * kH + col_of(i, lane) * 4;
      atomicAdd(w + 0, d[i] * f.x); atomicAdd(w + 1, d[i] * f.y); atomicAdd(w + 2, d[i] * f.z); atomicAdd(w + 3, d[i] * f.w);
    }
  }
  __syncthreads();
  flush(a.dout_g, sm + 0 * kH, kH); flush(a.dout_b, sm + 1 * kH, kH);
  flush(a.dimg_g, sm + 2 * kH, kH); flush(a.dimg_b, sm + 3 * kH, kH);
  flush(a.ddep_g, sm + 4 * kH, kH); flush(a.ddep_b, sm + 5 * kH, kH);
  flush(a.dloc_g, sm + 6 * kH, kH); flush(a.dloc_b, sm + 7 * kH, kH);
  flush(a.dloc_bias, sm + 8 * kH, kH); flush(a.dtok_emb1, sm + 9 * kH, kH);
  flush(a.dnav_emb, sm + 10 * kH, 2 * kH);
  flush(a.dloc_w, sm + 12 * kH, 4 * kH);
}

int pano_pack_bwd(const PanoPackBwdArgs& a, cudaStream_t stream) {
  if (a.rows <= 0) return ETP_OK;
  ETP_REQUIRE(a.dx && a.rgb_lin && a.loc_lin && a.sum_pre && a.stats && a.drgb_lin, "pano_pack_bwd: null argument");
  const size_t smem = 16 * kH * sizeof(float);
  static bool attr = false;
  if (!attr) {
    ETP_CHECK_CUDA(cudaFuncSetAttribute(pano_pack_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  int grid = grid_for(a.rows);
  if (grid > num_sms()) grid = num_sms();
  ETP_CHECK_CUDA(launch_pdl(pano_pack_bwd_kernel, dim3(grid), dim3(256), smem, stream, a));
  ETP_LAUNCHED();
  return ETP_OK;
}

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_txt_bwd_kernel(const float* __restrict__ dx, const int64_t* __restrict__ ids,
                                                             const float* __restrict__ sum_pre, const float* __restrict__ stats,
                                                             const float* __restrict__ gamma, int rows, int L, float* dword,
                                                             float* dpos, float* dtype0, float* dgamma, float* dbeta) {
  griddep_launch();  // PDL (common.cuh): let the next kernel get scheduled ...
  griddep_wait();    // ... and wait for the previous one before touching memory
  extern __shared__ float sm[];  // [3][768]: gamma, beta, type0
  zero_smem(sm, 3 * kH);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float g[24];
  ld24(gamma, lane, g);
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    float d[24], x[24], dgm[24], dbt[24];
    ld24(dx + static_cast<size_t>(row) * kH, lane, d);
    ld24(sum_pre + static_cast<size_t>(row) * kH, lane, x);
    ln_bwd_row(d, x, stats[static_cast<size_t>(row) * 2], stats[static_cast<size_t>(row) * 2 + 1], g, dgm, dbt);
    sacc(sm, lane, dgm);
    sacc(sm + kH, lane, dbt);
    sacc(sm + 2 * kH, lane, d);
    const int64_t id = ids[row];
    const int pos = row % L;
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      const int c = col_of(i, lane);
      if (dword && id != 0) atomicAdd(dword + id * kH + c, d[i]);  // padding_idx = 0 receives no gradient
      if (dpos) atomicAdd(dpos + static_cast<size_t>(pos) * kH + c, d[i]);
    }
  }
  __syncthreads();
  flush(dgamma, sm, kH);
  flush(dbeta, sm + kH, kH);
  flush(dtype0, sm + 2 * kH, kH);
}

int embed_txt_bwd(const float* dx, const int64_t* ids, const float* sum_pre, const float* stats, const float* gamma,
                  int B, int L, float* dword, float* dpos, float* dtype0, float* dgamma, float* dbeta,
                  cudaStream_t stream) {
  const int rows = B * L;
  if (rows <= 0) return ETP_OK;
  ETP_CHECK_CUDA(launch_pdl(embed_txt_bwd_kernel, dim3(grid_for(rows)), dim3(256), 3 * kH * sizeof(float), stream, dx, ids, sum_pre, stats, gamma, rows, L,
                                                                                 dword, dpos, dtype0, dgamma, dbeta));
  ETP_LAUNCHED();
  return ETP_OK;
}

// ---------------------------------------------------------------------------------------------------
// AdamW (torch.optim.AdamW, ss_trainer_ETP.py:213): p *= 1 - lr*wd ; m, v update ; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, bf16* __restrict__ pb, const float* __restrict__ g,
                                                     float* __restrict__ m, float* __restrict__ v, int64_t n4, float lr,
                                                     float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                                                     float gscale) {
  griddep_launch();  // PDL (common.cuh): let the next kernel get scheduled ...
  griddep_wait();    // ... and wait for the previous one before touching memory
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* pp = reinterpret_cast<float*>(&pv);
    const float* gg = reinterpret_cast<const float*>(&gv);
    float* mm = reinterpret_cast<float*>(&mv);
    float* vq = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = gg[k] * gscale;
      pp[k] *= 1.0f - lr * wd;
      mm[k] = b1 * mm[k] + (1.0f - b1) * gr;
      vq[k] = b2 * vq[k] + (1.0f - b2) * gr * gr;
      const float denom = sqrtf(vq[k]) / bc2_sqrt + eps;
      pp[k] -= (lr / bc1) * (mm[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (pb) reinterpret_cast<uint2*>(pb)[i] = make_uint2(pack_bf16x2(pp[0], pp[1]), pack_bf16x2(pp[2], pp[3]));
  }
}

int adamw_step(float* param, bf16* param_bf16, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
               float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
               cudaStream_t stream) {
  ETP_REQUIRE(n % 4 == 0, "adamw: element count must be a multiple of 4 (flat buffers are 64-element aligned)");
  ETP_REQUIRE(step >= 1, "adamw: step counts from 1");
  if (n == 0) return ETP_OK;
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > 16 * num_sms()) blocks = 16 * num_sms();
  ETP_CHECK_CUDA(launch_pdl(adamw_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, stream, param, param_bf16, grad, exp_avg, exp_avg_sq, n / 4, lr, beta1,
                                                             beta2, eps, weight_decay, bc1, sqrtf(bc2), grad_scale));
  ETP_LAUNCHED();
  return ETP_OK;
}

}  // namespace etp
