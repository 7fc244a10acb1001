// Token-packing kernels: panorama view tokens, graph-node tokens and the SAP-head tail.
// One warp per 768-wide token row (24 values per lane, 128-bit accesses, warp-shuffle LayerNorm).
//   pano_pack : vilmodel_cmt.py:695-711  (3 LayerNorms + nav-type / token-type embedding adds + LayerNorm;
//               the K=4 loc_linear is done here on CUDA cores, the K=512/128 linears come from the GEMM)
//   node_pack : vilmodel_cmt.py:728-730  (img_fts + step embedding + LN(Linear7(pos_fts)))
//   sap_tail  : vilmodel_cmt.py:654-658 (net.2 LayerNorm, net.4 Linear 768->1) + :742-744 (-inf masks)
#include "common.cuh"
#include "host.h"
#include "ops.h"

namespace etp {

constexpr int kH = 768;

ETP_DEVICE void ld24(const float* p, int lane, float (&v)[24]) {
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(p + (i * 32 + lane) * 4);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
ETP_DEVICE void st24(float* p, int lane, const float (&v)[24]) {
#pragma unroll
  for (int i = 0; i < 6; ++i)
    *reinterpret_cast<float4*>(p + (i * 32 + lane) * 4) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
ETP_DEVICE void st24_bf16(bf16* p, int lane, const float (&v)[24]) {
#pragma unroll
  for (int i = 0; i < 6; ++i)
    *reinterpret_cast<uint2*>(p + (i * 32 + lane) * 4) =
        make_uint2(pack_bf16x2(v[4 * i], v[4 * i + 1]), pack_bf16x2(v[4 * i + 2], v[4 * i + 3]));
}
ETP_DEVICE void stats24(const float (&v)[24], float eps, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) s += v[i];
  mean = warp_sum(s) * (1.0f / kH);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) { const float d = v[i] - mean; q += d * d; }
  rstd = rsqrtf(warp_sum(q) * (1.0f / kH) + eps);
}
// acc += LN(v) * g + b ; returns stats
ETP_DEVICE void ln_accum(const float (&v)[24], const float* g, const float* b, int lane, float eps, float (&acc)[24],
                         float& mean, float& rstd) {
  stats24(v, eps, mean, rstd);
  float gg[24], bb[24];
  ld24(g, lane, gg);
  ld24(b, lane, bb);
#pragma unroll
  for (int i = 0; i < 24; ++i) acc[i] += (v[i] - mean) * rstd * gg[i] + bb[i];
}

__global__ void __launch_bounds__(256) pano_pack_kernel(const PanoPackArgs a) {
  griddep_launch();  // PDL (common.cuh): let the next kernel get scheduled ...
  griddep_wait();    // ... and wait for the previous one before touching memory
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= a.rows) return;
  float acc[24], v[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) acc[i] = 0.f;
  float st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  ld24(a.rgb_lin + static_cast<size_t>(row) * kH, lane, v);
  ln_accum(v, a.img_g, a.img_b, lane, 1e-12f, acc, st[0], st[1]);
  if (a.dep_lin) {
    ld24(a.dep_lin + static_cast<size_t>(row) * kH, lane, v);
    ln_accum(v, a.dep_g, a.dep_b, lane, 1e-12f, acc, st[2], st[3]);
  }
  {
    const float4 f = *reinterpret_cast<const float4*>(a.loc_fts + static_cast<size_t>(row) * 4);
#pragma unroll
    for (int i = 0; i < 6; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = (i * 32 + lane) * 4 + j;
        const float4 w = __ldg(reinterpret_cast<const float4*>(a.loc_w + c * 4));
        v[4 * i + j] = f.x * w.x + f.y * w.y + f.z * w.z + f.w * w.w + __ldg(a.loc_b + c);
      }
    if (a.loc_lin) st24(a.loc_lin + static_cast<size_t>(row) * kH, lane, v);
    ln_accum(v, a.loc_g, a.loc_bb, lane, 1e-12f, acc, st[4], st[5]);
  }
  {
    const int nt = static_cast<int>(a.nav_types[row]);
    float e[24], t[24];
    ld24(a.nav_emb + nt * kH, lane, e);
    ld24(a.tok_emb1, lane, t);
#pragma unroll
    for (int i = 0; i < 24; ++i) acc[i] += e[i] + t[i];
  }
  if (a.sum_pre) st24(a.sum_pre + static_cast<size_t>(row) * kH, lane, acc);
#pragma unroll
  for (int i = 0; i < 24; ++i) v[i] = 0.f;
  ln_accum(acc, a.out_g, a.out_b, lane, 1e-12f, v, st[6], st[7]);
  drop_row24(Drop{a.drop.key, a.drop.thr, a.drop.scale}, row, lane, v);  // nn.Dropout after layer_norm (train mode)
  st24(a.x_f32 + static_cast<size_t>(row) * kH, lane, v);
  if (a.stats && lane < 8) {
    float s = st[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) s = (lane == i) ? st[i] : s;
    a.stats[static_cast<size_t>(row) * 8 + lane] = s;
  }
}

int pano_pack_fwd(const PanoPackArgs& a, cudaStream_t stream) {
  if (a.rows <= 0) return ETP_OK;
  ETP_REQUIRE(a.rgb_lin && a.loc_fts && a.nav_types && a.x_f32, "pano_pack: null argument");
  ETP_CHECK_CUDA(launch_pdl(pano_pack_kernel, dim3((a.rows + 7) / 8), dim3(256), 0, stream, a));
  ETP_LAUNCHED();
  return ETP_OK;
}

__global__ void __launch_bounds__(256) node_pack_kernel(const NodePackArgs a) {
  griddep_launch();  // PDL (common.cuh): let the next kernel get scheduled ...
  griddep_wait();    // ... and wait for the previous one before touching memory
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= a.rows) return;
  float f[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) f[j] = a.pos_fts[static_cast<size_t>(row) * 7 + j];
  float v[24], acc[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) {
    const int c = ((i >> 2) * 32 + lane) * 4 + (i & 3);
    float s = __ldg(a.pos_b + c);
#pragma unroll
    for (int j = 0; j < 7; ++j) s += f[j] * __ldg(a.pos_w + c * 7 + j);
    v[i] = s;
  }
  if (a.pos_lin) st24(a.pos_lin + static_cast<size_t>(row) * kH, lane, v);
  ld24(a.img_fts + static_cast<size_t>(row) * kH, lane, acc);
  {
    float e[24];
    ld24(a.step_emb + static_cast<size_t>(a.step_ids[row]) * kH, lane, e);
#pragma unroll
    for (int i = 0; i < 24; ++i) acc[i] += e[i];
  }
  float mean, rstd;
  ln_accum(v, a.pos_g, a.pos_bb, lane, 1e-12f, acc, mean, rstd);
  st24(a.x_f32 + static_cast<size_t>(row) * kH, lane, acc);
  if (a.x_bf16) st24_bf16(a.x_bf16 + static_cast<size_t>(row) * kH, lane, acc);
  if (a.stats && lane == 0) {
    a.stats[static_cast<size_t>(row) * 2] = mean;
    a.stats[static_cast<size_t>(row) * 2 + 1] = rstd;
  }
}

int node_pack_fwd(const NodePackArgs& a, cudaStream_t stream) {
  if (a.rows <= 0) return ETP_OK;
  ETP_REQUIRE(a.img_fts && a.step_ids && a.pos_fts && a.x_f32, "node_pack: null argument");
  ETP_CHECK_CUDA(launch_pdl(node_pack_kernel, dim3((a.rows + 7) / 8), dim3(256), 0, stream, a));
  ETP_LAUNCHED();
  return ETP_OK;
}

__global__ void __launch_bounds__(256) sap_tail_kernel(const float* __restrict__ relu_out, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, const float* __restrict__ w4,
                                                        const float* __restrict__ b4, const uint8_t* __restrict__ visited,
                                                        const uint8_t* __restrict__ valid, int rows,
                                                        float* __restrict__ logits, float* __restrict__ mean_out,
                                                        float* __restrict__ rstd_out, const Drop drop) {
  griddep_launch();  // PDL (common.cuh): let the next kernel get scheduled ...
  griddep_wait();    // ... and wait for the previous one before touching memory
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float v[24], h[24], w[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) h[i] = 0.f;
  ld24(relu_out + static_cast<size_t>(row) * kH, lane, v);
  float mean, rstd;
  ln_accum(v, gamma, beta, lane, 1e-12f, h, mean, rstd);
  drop_row24(drop, row, lane, h);  // NextActionPrediction's Dropout between LayerNorm and the last Linear
  ld24(w4, lane, w);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) s += h[i] * w[i];
  s = warp_sum(s) + b4[0];
  if (lane == 0) {
    const bool dead = (visited && visited[row]) || (valid && !valid[row]);
    logits[row] = dead ? -INFINITY : s;
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

int sap_tail_fwd(const float* relu_out, const float* gamma, const float* beta, const float* w4, const float* b4,
                 const uint8_t* visited, const uint8_t* valid, int rows, int H, float* logits, float* mean,
                 float* rstd, cudaStream_t stream, DropHost drop) {
  ETP_REQUIRE(H == kH, "sap_tail: hidden size must be 768");
  if (rows <= 0) return ETP_OK;
  ETP_CHECK_CUDA(launch_pdl(sap_tail_kernel, dim3((rows + 7) / 8), dim3(256), 0, stream, relu_out, gamma, beta, w4, b4, visited, valid, rows, logits, mean,
                                                      rstd, Drop{drop.key, drop.thr, drop.scale}));
  ETP_LAUNCHED();
  return ETP_OK;
}

}  // namespace etp

namespace etp {

__global__ void __launch_bounds__(256) embed_txt_kernel(const int64_t* __restrict__ ids, const float* __restrict__ word_emb,
                                                         const float* __restrict__ pos_emb, const float* __restrict__ type_emb0,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         float eps, int rows, int L, float* __restrict__ x_f32,
                                                         bf16* __restrict__ x_bf16, float* __restrict__ sum_pre,
                                                         float* __restrict__ stats, const Drop drop) {
  griddep_launch();  // PDL (common.cuh): let the next kernel get scheduled ...
  griddep_wait();    // ... and wait for the previous one before touching memory
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  float w[24], p[24], t[24], v[24];
  ld24(word_emb + static_cast<size_t>(ids[row]) * kH, lane, w);
  ld24(pos_emb + static_cast<size_t>(row % L) * kH, lane, p);
  ld24(type_emb0, lane, t);
#pragma unroll
  for (int i = 0; i < 24; ++i) { w[i] += p[i] + t[i]; v[i] = 0.f; }
  if (sum_pre) st24(sum_pre + static_cast<size_t>(row) * kH, lane, w);
  float mean, rstd;
  ln_accum(w, gamma, beta, lane, eps, v, mean, rstd);
  drop_row24(drop, row, lane, v);  // BertEmbeddings.dropout (train mode)
  st24(x_f32 + static_cast<size_t>(row) * kH, lane, v);
  if (x_bf16) st24_bf16(x_bf16 + static_cast<size_t>(row) * kH, lane, v);
  if (stats && lane == 0) {
    stats[static_cast<size_t>(row) * 2] = mean;
    stats[static_cast<size_t>(row) * 2 + 1] = rstd;
  }
}

int embed_txt_fwd(const int64_t* ids, const float* word_emb, const float* pos_emb, const float* type_emb0,
                  const float* gamma, const float* beta, float eps, int B, int L, float* x_f32, bf16* x_bf16,
                  float* sum_pre, float* stats, cudaStream_t stream, DropHost drop) {
  const int rows = B * L;
  if (rows <= 0) return ETP_OK;
  ETP_CHECK_CUDA(launch_pdl(embed_txt_kernel, dim3((rows + 7) / 8), dim3(256), 0, stream, ids, word_emb, pos_emb, type_emb0, gamma, beta, eps, rows, L,
                                                       x_f32, x_bf16, sum_pre, stats, Drop{drop.key, drop.thr, drop.scale}));
  ETP_LAUNCHED();
  return ETP_OK;
}

__global__ void seq_mask_kernel(const int64_t* __restrict__ lens, int B, int V, uint8_t* __restrict__ mask) {
  griddep_launch();  // PDL (common.cuh): let the next kernel get scheduled ...
  griddep_wait();    // ... and wait for the previous one before touching memory
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * V) mask[i] = (i % V) < lens[i / V] ? 1 : 0;
}
int seq_mask(const int64_t* lens, int B, int V, uint8_t* mask, cudaStream_t stream) {
  if (B * V <= 0) return ETP_OK;
  ETP_CHECK_CUDA(launch_pdl(seq_mask_kernel, dim3((B * V + 255) / 256), dim3(256), 0, stream, lens, B, V, mask));
  ETP_LAUNCHED();
  return ETP_OK;
}

// ---------------------------------------------------------------------------------------------------
// Caller-side loss of one step, fused (ss_trainer_ETP.py:879-900): softmax over the node logits (-inf entries are
// masked nodes), cross-entropy summed over the batch with ignore_index, its gradient w.r.t. the logits, and the
// greedy action.  One warp per episode row.
//   loss_sum += sum_b CE(logits[b], labels[b])          (rows with labels[b] == ignore_index contribute nothing)
//   dlogits[b, n] = grad_scale * (softmax[b, n] - [n == labels[b]])      (0 for ignored rows and -inf entries)
//   probs[b, n] = softmax[b, n]  (optional; the trainer samples from it),  argmax[b] = first maximal logit
__global__ void __launch_bounds__(256) step_loss_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int B,
                                                         int N, int64_t ignore_index, float grad_scale,
                                                         float* __restrict__ loss_sum, float* __restrict__ dlogits,
                                                         float* __restrict__ probs, int64_t* __restrict__ argmax) {
  griddep_launch();  // PDL (common.cuh): let the next kernel get scheduled ...
  griddep_wait();    // ... and wait for the previous one before touching memory
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* x = logits + static_cast<size_t>(b) * N;
  float mx = -INFINITY;
  int arg = 0x7fffffff;
  for (int n = lane; n < N; n += 32) {
    const float v = x[n];
    if (v > mx) { mx = v; arg = n; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
  }
  float se = 0.f;
  for (int n = lane; n < N; n += 32) se += __expf(x[n] - mx);
  se = warp_sum(se);
  const float inv = 1.0f / se;
  const int64_t lab = labels ? labels[b] : ignore_index;
  const bool live = lab != ignore_index && lab >= 0 && lab < N;
  for (int n = lane; n < N; n += 32) {
    const float pr = __expf(x[n] - mx) * inv;
    if (probs) probs[static_cast<size_t>(b) * N + n] = pr;
    if (dlogits) dlogits[static_cast<size_t>(b) * N + n] = live ? grad_scale * (pr - (n == lab ? 1.0f : 0.0f)) : 0.0f;
  }
  if (lane == 0) {
    if (argmax) argmax[b] = arg == 0x7fffffff ? 0 : arg;
    if (live && loss_sum) atomicAdd(loss_sum, -(x[lab] - mx - __logf(se)));
  }
}

int step_loss(const float* logits, const int64_t* labels, int B, int N, int64_t ignore_index, float grad_scale,
              float* loss_sum, float* dlogits, float* probs, int64_t* argmax, cudaStream_t stream) {
  ETP_REQUIRE(logits != nullptr && B > 0 && N > 0, "step_loss: bad arguments");
  ETP_CHECK_CUDA(launch_pdl(step_loss_kernel, dim3((B + 7) / 8), dim3(256), 0, stream, logits, labels, B, N, ignore_index,
                            grad_scale, loss_sum, dlogits, probs, argmax));
  ETP_LAUNCHED();
  return ETP_OK;
}

}  // namespace etp
