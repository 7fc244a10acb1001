// Backward of the token-packing kernels, the SAP-head tail and the text embedding (pack.cu), plus AdamW.
// Same mapping as the forward: one warp per 768-wide row, 24 values per lane.  Parameter gradients are accumulated
// as per-lane register partials over a grid-stride loop of rows and reduced once per CTA (layernorm_bwd pattern),
// or, where only a few hundred rows exist, added straight to global memory with 128-bit vector reductions.
// (Shared-memory float atomics are CAS loops: with eight warps adding to the same 768 addresses they cost ~20 us
// per ROW in the first version of these kernels.)
#include <atomic>

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
// one row's contribution straight to global memory: six 128-bit reductions per lane (512 B per warp instruction)
ETP_DEVICE void gred24(float* dst, int lane, const float (&v)[24]) {
  if (dst == nullptr) return;
#pragma unroll
  for (int i = 0; i < 6; ++i)
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + (i * 32 + lane) * 4), "f"(v[4 * i]),
                 "f"(v[4 * i + 1]), "f"(v[4 * i + 2]), "f"(v[4 * i + 3])
                 : "memory");
}
// CTA reduction of per-lane register partials (one 24-vector per warp) through red[8][768], then one atomicAdd per
// column per CTA (the layernorm_bwd pattern).  Every thread of the 256-thread CTA must call it.
ETP_DEVICE void cta_reduce_add(float (*red)[kH], int warp, int lane, const float (&v)[24], float* dst) {
  if (dst == nullptr) return;  // uniform across the CTA
#pragma unroll
  for (int i = 0; i < 24; ++i) red[warp][col_of(i, lane)] = v[i];
  __syncthreads();
  for (int c = threadIdx.x; c < kH; c += 256) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][c];
    atomicAdd(dst + c, t);
  }
  __syncthreads();
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
                                                            float* dgamma, float* dbeta, float* dw4, float* db4, const Drop drop) {
  griddep_launch();  // PDL (common.cuh): let the next kernel get scheduled ...
  griddep_wait();    // ... and wait for the previous one before touching memory
  __shared__ float red[8][kH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float g[24], bt[24], w[24];
  ld24(gamma, lane, g);
  ld24(beta, lane, bt);
  ld24(w4, lane, w);
  // parameter-gradient partials of this lane over the warp's rows (registers), reduced once at the end
  float ag[24], ab[24], aw[24], ab4 = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) { ag[i] = 0.f; ab[i] = 0.f; aw[i] = 0.f; }
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    const bool dead = (visited && visited[row]) || (valid && !valid[row]);
    const float dl = dead ? 0.f : dlogits[row];
    float r[24], dh[24], x[24], dgm[24], dbt[24];
    ld24(relu_out + static_cast<size_t>(row) * kH, lane, r);
    const float mu = mean[row], rs = rstd[row];
    float hm[24];  // the forward's dropout multipliers of this row (all 1 when off)
#pragma unroll
    for (int i = 0; i < 24; ++i) hm[i] = 1.0f;
    drop_row24(drop, row, lane, hm);
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      x[i] = r[i];
      dh[i] = dl * w[i] * hm[i];
      aw[i] = fmaf(dl, ((r[i] - mu) * rs * g[i] + bt[i]) * hm[i], aw[i]);
    }
    ln_bwd_row(dh, x, mu, rs, g, dgm, dbt);
#pragma unroll
    for (int i = 0; i < 24; ++i) {
      ag[i] += dgm[i];
      ab[i] += dbt[i];
      dh[i] = r[i] > 0.f ? dh[i] : 0.f;
    }
    st24_bf16(dpre + static_cast<size_t>(row) * kH, lane, dh);
    ab4 += dl;
  }
  cta_reduce_add(red, warp, lane, ag, dgamma);
  cta_reduce_add(red, warp, lane, ab, dbeta);
  cta_reduce_add(red, warp, lane, aw, dw4);
  if (db4 != nullptr && lane == 0 && ab4 != 0.f) atomicAdd(db4, ab4);
}

int sap_tail_bwd(const float* dlogits, const float* relu_out, const float* gamma, const float* beta, const float* w4,
                 const float* mean, const float* rstd, const uint8_t* visited, const uint8_t* valid, int rows,
                 bf16* dpre_bf16, float* dgamma, float* dbeta, float* dw4, float* db4, cudaStream_t stream, DropHost drop) {
  if (rows <= 0) return ETP_OK;
  ETP_CHECK_CUDA(launch_pdl(sap_tail_bwd_kernel, dim3(grid_for(rows)), dim3(256), 0, stream, 
      dlogits, relu_out, gamma, beta, w4, mean, rstd, visited, valid, rows, dpre_bf16, dgamma, dbeta, dw4, db4,
      Drop{drop.key, drop.thr, drop.scale}));
  ETP_LAUNCHED();
  return ETP_OK;
}

// ---------------------------------------------------------------------------------------------------
// x0 = img_fts + E_step[ids] + LN(pos_fts.W^T + b): LayerNorm backward per row (one warp per row); gamma / beta /
// bias / step-0 gradients as per-lane register partials; the [768,7] weight gradient d^T.pos_fts by staging the 8 rows
// of an iteration in shared memory and letting each thread own 3 columns (21 fp32 accumulators); rows with a
// non-zero step id (visited nodes, a minority) add straight to their embedding row with vector reductions.
__global__ void __launch_bounds__(256) node_pack_bwd_kernel(const float* __restrict__ dx, const int64_t* __restrict__ step_ids,
                                                             const float* __restrict__ pos_fts, const float* __restrict__ pos_lin,
                                                             const float* __restrict__ stats, const float* __restrict__ pos_g,
                                                             int rows, float* dstep_emb, float* dpos_w, float* dpos_b,
                                                             float* dpos_g, float* dpos_bb) {
  griddep_launch();  // PDL (common.cuh): let the next kernel get scheduled ...
  griddep_wait();    // ... and wait for the previous one before touching memory
  __shared__ float red[8][kH];  // row staging for the weight gradient, then the CTA reduction buffer
  __shared__ float sf[8][8];    // pos_fts of the 8 rows of this iteration
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float g[24];
  ld24(pos_g, lane, g);
  float ag[24], ab[24], abias[24], astep0[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) { ag[i] = 0.f; ab[i] = 0.f; abias[i] = 0.f; astep0[i] = 0.f; }
  float aw[3][7];  // thread-owned columns threadIdx.x, +256, +512 of dpos_w
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int j = 0; j < 7; ++j) aw[c][j] = 0.f;
  const int iters = (rows + gridDim.x * 8 - 1) / (gridDim.x * 8);
  for (int it = 0; it < iters; ++it) {
    const int row = (it * gridDim.x + blockIdx.x) * 8 + warp;
    float d[24];
    if (row < rows) {
      float x[24], dgm[24], dbt[24];
      ld24(dx + static_cast<size_t>(row) * kH, lane, d);
      const int64_t id = step_ids[row];
      if (dstep_emb) {
        if (id == 0) {
#pragma unroll
          for (int i = 0; i < 24; ++i) astep0[i] += d[i];
        } else {
          gred24(dstep_emb + id * kH, lane, d);
        }
      }
      ld24(pos_lin + static_cast<size_t>(row) * kH, lane, x);
      ln_bwd_row(d, x, stats[static_cast<size_t>(row) * 2], stats[static_cast<size_t>(row) * 2 + 1], g, dgm, dbt);
#pragma unroll
      for (int i = 0; i < 24; ++i) { ag[i] += dgm[i]; ab[i] += dbt[i]; abias[i] += d[i]; }
      if (lane < 7) sf[warp][lane] = pos_fts[static_cast<size_t>(row) * 7 + lane];
    } else {
#pragma unroll
      for (int i = 0; i < 24; ++i) d[i] = 0.f;
      if (lane < 7) sf[warp][lane] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 24; ++i) red[warp][col_of(i, lane)] = d[i];
    __syncthreads();
    if (dpos_w) {
#pragma unroll
      for (int r8 = 0; r8 < 8; ++r8) {
        float f[7];
#pragma unroll
        for (int j = 0; j < 7; ++j) f[j] = sf[r8][j];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float dv = red[r8][threadIdx.x + 256 * c];
#pragma unroll
          for (int j = 0; j < 7; ++j) aw[c][j] = fmaf(dv, f[j], aw[c][j]);
        }
      }
    }
    __syncthreads();
  }
  cta_reduce_add(red, warp, lane, ag, dpos_g);
  cta_reduce_add(red, warp, lane, ab, dpos_bb);
  cta_reduce_add(red, warp, lane, abias, dpos_b);
  cta_reduce_add(red, warp, lane, astep0, dstep_emb);  // row 0 of the step-embedding table
  if (dpos_w) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int j = 0; j < 7; ++j) atomicAdd(dpos_w + (threadIdx.x + 256 * c) * 7 + j, aw[c][j]);
  }
}

int node_pack_bwd(const float* dx, const int64_t* step_ids, const float* pos_fts, const float* pos_lin,
                  const float* stats, const float* pos_g, int rows, float* dstep_emb, float* dpos_w, float* dpos_b,
                  float* dpos_g, float* dpos_bb, cudaStream_t stream) {
  if (rows <= 0) return ETP_OK;
  int grid = (rows + 7) / 8;
  if (grid > num_sms()) grid = num_sms();
  ETP_CHECK_CUDA(launch_pdl(node_pack_bwd_kernel, dim3(grid), dim3(256), 0, stream, dx, step_ids, pos_fts, pos_lin, stats, pos_g, rows, dstep_emb, dpos_w,
                                                    dpos_b, dpos_g, dpos_bb));
  ETP_LAUNCHED();
  return ETP_OK;
}

// ---------------------------------------------------------------------------------------------------
// Backward of the view-token packing (4 LayerNorms, nav-type / token-type embeddings, the K=4 location Linear).
// Only B*V rows (<= a few thousand): one warp per row and every parameter gradient goes straight to global memory
// with 128-bit vector reductions (512 B per warp instruction) — no shared-memory accumulation, no CTA reduction.
__global__ void __launch_bounds__(256) pano_pack_bwd_kernel(const PanoPackBwdArgs a) {
  griddep_launch();  // PDL (common.cuh): let the next kernel get scheduled ...
  griddep_wait();    // ... and wait for the previous one before touching memory
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float og[24];
  ld24(a.out_g, lane, og);
  for (int row = blockIdx.x * 8 + warp; row < a.rows; row += gridDim.x * 8) {
    const float* st = a.stats + static_cast<size_t>(row) * 8;
    float d[24], x[24], dgm[24], dbt[24], g[24];
    ld24(a.dx + static_cast<size_t>(row) * kH, lane, d);
    drop_row24(Drop{a.drop.key, a.drop.thr, a.drop.scale}, row, lane, d);  // back through the forward's dropout
    ld24(a.sum_pre + static_cast<size_t>(row) * kH, lane, x);
    ln_bwd_row(d, x, st[6], st[7], og, dgm, dbt);   // d = dsum
    gred24(a.dout_g, lane, dgm);
    gred24(a.dout_b, lane, dbt);
    gred24(a.dtok_emb1, lane, d);
    if (a.dnav_emb) gred24(a.dnav_emb + static_cast<int>(a.nav_types[row]) * kH, lane, d);
    float t[24];
    // img branch
#pragma unroll
    for (int i = 0; i < 24; ++i) t[i] = d[i];
    ld24(a.rgb_lin + static_cast<size_t>(row) * kH, lane, x);
    ld24(a.img_g, lane, g);
    ln_bwd_row(t, x, st[0], st[1], g, dgm, dbt);
    gred24(a.dimg_g, lane, dgm);
    gred24(a.dimg_b, lane, dbt);
    st24_bf16(a.drgb_lin + static_cast<size_t>(row) * kH, lane, t);
    if (a.dep_lin) {
#pragma unroll
      for (int i = 0; i < 24; ++i) t[i] = d[i];
      ld24(a.dep_lin + static_cast<size_t>(row) * kH, lane, x);
      ld24(a.dep_g, lane, g);
      ln_bwd_row(t, x, st[2], st[3], g, dgm, dbt);
      gred24(a.ddep_g, lane, dgm);
      gred24(a.ddep_b, lane, dbt);
      st24_bf16(a.ddep_lin + static_cast<size_t>(row) * kH, lane, t);
    }
    ld24(a.loc_lin + static_cast<size_t>(row) * kH, lane, x);
    ld24(a.loc_g, lane, g);
    ln_bwd_row(d, x, st[4], st[5], g, dgm, dbt);   // d = dloc_lin
    gred24(a.dloc_g, lane, dgm);
    gred24(a.dloc_b, lane, dbt);
    gred24(a.dloc_bias, lane, d);
    if (a.dloc_w) {
      // dW[c, 0..3] += d[c] * loc_fts[row, 0..3]: the 4 columns a lane holds per float4 group are 16 consecutive floats
      const float4 f = *reinterpret_cast<const float4*>(a.loc_fts + static_cast<size_t>(row) * 4);
#pragma unroll
      for (int i = 0; i < 24; ++i)
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.dloc_w + col_of(i, lane) * 4), "f"(d[i] * f.x),
                     "f"(d[i] * f.y), "f"(d[i] * f.z), "f"(d[i] * f.w)
                     : "memory");
    }
  }
}

int pano_pack_bwd(const PanoPackBwdArgs& a, cudaStream_t stream) {
  if (a.rows <= 0) return ETP_OK;
  ETP_REQUIRE(a.dx && a.rgb_lin && a.loc_lin && a.sum_pre && a.stats && a.drgb_lin, "pano_pack_bwd: null argument");
  const int grid = (a.rows + 7) / 8;
  ETP_CHECK_CUDA(launch_pdl(pano_pack_bwd_kernel, dim3(grid), dim3(256), 0, stream, a));
  ETP_LAUNCHED();
  return ETP_OK;
}

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_txt_bwd_kernel(const float* __restrict__ dx, const int64_t* __restrict__ ids,
                                                             const float* __restrict__ sum_pre, const float* __restrict__ stats,
                                                             const float* __restrict__ gamma, int rows, int L, float* dword,
                                                             float* dpos, float* dtype0, float* dgamma, float* dbeta,
                                                             const Drop drop) {
  griddep_launch();  // PDL (common.cuh): let the next kernel get scheduled ...
  griddep_wait();    // ... and wait for the previous one before touching memory
  __shared__ float red[8][kH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float g[24];
  ld24(gamma, lane, g);
  float ag[24], ab[24], at0[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) { ag[i] = 0.f; ab[i] = 0.f; at0[i] = 0.f; }
  for (int row = blockIdx.x * 8 + warp; row < rows; row += gridDim.x * 8) {
    float d[24], x[24], dgm[24], dbt[24];
    ld24(dx + static_cast<size_t>(row) * kH, lane, d);
    drop_row24(drop, row, lane, d);  // back through the forward's dropout
    ld24(sum_pre + static_cast<size_t>(row) * kH, lane, x);
    ln_bwd_row(d, x, stats[static_cast<size_t>(row) * 2], stats[static_cast<size_t>(row) * 2 + 1], g, dgm, dbt);
#pragma unroll
    for (int i = 0; i < 24; ++i) { ag[i] += dgm[i]; ab[i] += dbt[i]; at0[i] += d[i]; }
    const int64_t id = ids[row];
    const int pos = row % L;
    if (dword && id != 0) gred24(dword + id * kH, lane, d);  // padding_idx = 0 receives no gradient
    if (dpos) gred24(dpos + static_cast<size_t>(pos) * kH, lane, d);
  }
  cta_reduce_add(red, warp, lane, ag, dgamma);
  cta_reduce_add(red, warp, lane, ab, dbeta);
  cta_reduce_add(red, warp, lane, at0, dtype0);
}

int embed_txt_bwd(const float* dx, const int64_t* ids, const float* sum_pre, const float* stats, const float* gamma,
                  int B, int L, float* dword, float* dpos, float* dtype0, float* dgamma, float* dbeta,
                  cudaStream_t stream, DropHost drop) {
  const int rows = B * L;
  if (rows <= 0) return ETP_OK;
  ETP_CHECK_CUDA(launch_pdl(embed_txt_bwd_kernel, dim3(grid_for(rows)), dim3(256), 0, stream, dx, ids, sum_pre, stats, gamma, rows, L,
                                                                                 dword, dpos, dtype0, dgamma, dbeta,
                                                                                 Drop{drop.key, drop.thr, drop.scale}));
  ETP_LAUNCHED();
  return ETP_OK;
}

// ---------------------------------------------------------------------------------------------------
// AdamW (torch.optim.AdamW, ss_trainer_ETP.py:213): p *= 1 - lr*wd ; m, v update ; p -= lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
// flags (optional): one byte per 64-element block of the flat layout (every tensor starts on a 64-element boundary):
// bit 0 = the block belongs to a trainable parameter (torch's AdamW skips parameters without a gradient), bit 1 = weight
// decay applies (the reference's pre-training optimizer has a no-decay group for biases and LayerNorm, optim/misc.py:14-20).
// normsq (optional): device scalar holding sum g^2 over the trainable blocks; with max_norm > 0 the gradient is scaled by
// min(1, max_norm / (gscale * sqrt(normsq) + 1e-6)) like torch.nn.utils.clip_grad_norm_ (train_r2r.py:279-284).
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, bf16* __restrict__ pb, const float* __restrict__ g,
                                                     float* __restrict__ m, float* __restrict__ v, int64_t n4, float lr,
                                                     float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                                                     float gscale, const uint8_t* __restrict__ flags,
                                                     const float* __restrict__ normsq, float max_norm) {
  griddep_launch();  // PDL (common.cuh): let the next kernel get scheduled ...
  griddep_wait();    // ... and wait for the previous one before touching memory
  if (normsq != nullptr && max_norm > 0.0f) {
    const float total = gscale * sqrtf(__ldg(normsq));
    gscale *= fminf(1.0f, max_norm / (total + 1e-6f));
  }
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint32_t fl = flags ? flags[i >> 4] : 3u;
    if (!(fl & 1u)) continue;
    const float decay = (fl & 2u) ? 1.0f - lr * wd : 1.0f;
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
      pp[k] *= decay;
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

static std::atomic<int> g_adamw_ctas_per_sm{16};
void set_adamw_ctas_per_sm(int n) { g_adamw_ctas_per_sm.store(n < 1 ? 1 : (n > 16 ? 16 : n)); }

int adamw_step(float* param, bf16* param_bf16, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
               float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
               cudaStream_t stream, const uint8_t* flags, const float* normsq, float max_norm) {
  ETP_REQUIRE(n % 4 == 0, "adamw: element count must be a multiple of 4 (flat buffers are 64-element aligned)");
  ETP_REQUIRE(step >= 1, "adamw: step counts from 1");
  if (n == 0) return ETP_OK;
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  int64_t blocks = (n / 4 + 255) / 256;
  // default: 16 CTAs per SM (the update alone at the HBM roofline).  A host that runs the update on a side stream UNDER
  // the backward pass (PlannerTrainer) asks for a small grid (etp_set_adamw_ctas_per_sm): one grid-striding CTA per SM
  // streams at a fraction of the bandwidth but leaves the registers / thread slots the persistent GEMM CTAs need.
  const int per_sm = g_adamw_ctas_per_sm.load(std::memory_order_relaxed);
  if (blocks > static_cast<int64_t>(per_sm) * num_sms()) blocks = static_cast<int64_t>(per_sm) * num_sms();
  ETP_CHECK_CUDA(launch_pdl(adamw_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, stream, param, param_bf16, grad, exp_avg, exp_avg_sq, n / 4, lr, beta1,
                                                             beta2, eps, weight_decay, bc1, sqrtf(bc2), grad_scale, flags, normsq, max_norm));
  ETP_LAUNCHED();
  return ETP_OK;
}

// out[0] += sum of g^2 over the trainable 64-element blocks (the squared global gradient norm of clip_grad_norm_)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n4, const uint8_t* __restrict__ flags,
                                                     float* __restrict__ out) {
  griddep_launch();
  griddep_wait();
  float acc = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (flags && !(flags[i >> 4] & 1u)) continue;
    const float4 t = reinterpret_cast<const float4*>(g)[i];
    acc += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
  }
  acc = warp_sum(acc);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += part[k];
    atomicAdd(out, s);
  }
}

int grad_sumsq(const float* grad, int64_t n, const uint8_t* flags, float* out, cudaStream_t stream) {
  ETP_REQUIRE(n % 4 == 0 && grad && out, "grad_sumsq: bad argument");
  if (n == 0) return ETP_OK;
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  ETP_CHECK_CUDA(launch_pdl(sumsq_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, stream, grad, n / 4, flags, out));
  ETP_LAUNCHED();
  return ETP_OK;
}

}  // namespace etp
