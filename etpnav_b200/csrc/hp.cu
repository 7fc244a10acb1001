// High-precision forward mode ("precision = high"): the reference's eval / inference path runs in fp32
// (ss_trainer_ETP.py:513-756 has no autocast), and BASELINE.json's north_star states the parity band
// rtol 1e-3 / atol 1e-4, which bf16 operands (2^-9 relative rounding) cannot meet.  This mode keeps every
// contraction on the tcgen05 tensor cores and recovers fp32-class accuracy by SPLIT-bf16 x3 products:
//
//     a = a_hi + a_lo (+ O(2^-17 a)),  a_hi = bf16(a), a_lo = bf16(a - a_hi)
//     a.b  ~=  a_hi.b_hi + a_lo.b_hi + a_hi.b_lo              (the lo.lo term is 2^-18 relative)
//
// realised with NO change to the GEMM kernel: the three products are one GEMM over a K axis three times as long,
//     A' = [ A_hi | A_lo | A_hi ]  (rows x 3K),   B' = [ B_hi | B_hi | B_lo ]  (N x 3K),   D = A'.B'^T
// (fp32 accumulation in TMEM as before).  split3_kernel writes those layouts from fp32 activations / master
// weights; every activation of this mode stays fp32 in HBM, the epilogues (bias, GELU, ReLU, residual) and
// LayerNorm are the fp32 code the bf16 mode already uses.  Attention (softmax(QK^T)V, 4 % of the step's FLOPs)
// runs in fp32 on the CUDA cores (attention_f32_kernel): scores, softmax and P.V never see a bf16 rounding.
// Inference only: no activation record is written and there is no backward in this mode.
// Replaces, like the bf16 sequencing in planner.cu: vilmodel_cmt.py:684-750, common/transformer.py:170-182.
#include "../../include/etpnav_b200.h"
#include "common.cuh"
#include "planner.h"

namespace etp {

#define ETP_TRY(expr)              \
  do {                             \
    int _rc = (expr);              \
    if (_rc != ETP_OK) return _rc; \
  } while (0)

// ---------------------------------------------------------------------------------------------------
// fp32 [rows, K] (pitch ldx) -> bf16 [rows, 3K]:  form 0 (A operand) hi | lo | hi,  form 1 (B operand) hi | hi | lo
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ x, bf16* __restrict__ y, int64_t rows, int K,
                                                     int64_t ldx, int form) {
  griddep_launch();
  griddep_wait();
  const int k4 = K >> 2;  // K % 4 == 0
  const int64_t total = rows * k4;
  const int lo_slot = form == 0 ? 1 : 2, hi2_slot = form == 0 ? 2 : 1;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / k4;
    const int c = static_cast<int>(i - r * k4) * 4;
    const float4 t = *reinterpret_cast<const float4*>(x + r * ldx + c);
    const float v[4] = {t.x, t.y, t.z, t.w};
    float hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      hi[j] = __bfloat162float(__float2bfloat16_rn(v[j]));
      lo[j] = v[j] - hi[j];  // exact in fp32
    }
    const uint2 h = make_uint2(pack_bf16x2(hi[0], hi[1]), pack_bf16x2(hi[2], hi[3]));
    const uint2 l = make_uint2(pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]));
    bf16* row = y + r * (3 * static_cast<int64_t>(K));
    *reinterpret_cast<uint2*>(row + c) = h;
    *reinterpret_cast<uint2*>(row + static_cast<int64_t>(lo_slot) * K + c) = l;
    *reinterpret_cast<uint2*>(row + static_cast<int64_t>(hi2_slot) * K + c) = h;
  }
}

int split3(const float* x, bf16* y, int64_t rows, int K, int64_t ldx, int form, cudaStream_t s) {
  if (rows <= 0) return ETP_OK;
  ETP_REQUIRE(x && y, "split3: null argument");
  ETP_REQUIRE(K > 0 && K % 4 == 0 && ldx % 4 == 0 && (form == 0 || form == 1), "split3: K and pitch must be multiples of 4");
  ETP_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0, "split3: alignment");
  int64_t blocks = (rows * (K / 4) + 255) / 256;
  if (blocks > 16 * static_cast<int64_t>(num_sms())) blocks = 16 * num_sms();
  ETP_CHECK_CUDA(launch_pdl(split3_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, s, x, y, rows, K, ldx, form));
  ETP_LAUNCHED();
  return ETP_OK;
}

// ---------------------------------------------------------------------------------------------------
// fp32 attention, head dim 64:  out = softmax(scale * q.k^T + bias) . v  with the bias of ops.h: AttnArgs.
// One CTA = 16 query rows of one (batch, head), 8 warps x 2 rows; keys in chunks of 64 staged in shared memory
// (K chunk padded to 65 floats per row: lane = key reads are conflict-free; V chunk: lane = output column pair),
// online softmax across chunks (exact: the running maximum only rescales).
// ---------------------------------------------------------------------------------------------------
namespace {
constexpr int kFD = 64, kFChunk = 64, kFRows = 16, kFKPitch = 65;

struct AttnF32 {
  int B, heads, Sq, Sk;
  const float *q, *k, *v;
  int ldq, ldk, ldv;
  float scale;
  const uint8_t* key_valid;
  float mask_value;
  const float* pair;
  const float *pair_w_dev, *pair_b_dev;
  float pair_w, pair_b;
  float* out;
  int ldo;
};

__global__ void __launch_bounds__(256) attention_f32_kernel(const AttnF32 a) {
  griddep_launch();
  griddep_wait();
  __shared__ float Ks[kFChunk * kFKPitch];
  __shared__ float Vs[kFChunk * kFD];
  __shared__ float kb[kFChunk];
  __shared__ float qs[kFRows * kFD];
  __shared__ float ps[kFRows * kFChunk];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kFRows;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float pw = a.pair_w_dev ? __ldg(a.pair_w_dev) : a.pair_w;
  const float pb = a.pair_b_dev ? __ldg(a.pair_b_dev) : a.pair_b;
  // the CTA's query rows (pre-scaled)
  for (int i = threadIdx.x; i < kFRows * kFD; i += 256) {
    const int r = i >> 6, d = i & 63;
    const int q = q0 + r;
    qs[i] = q < a.Sq ? a.q[(static_cast<size_t>(b) * a.Sq + q) * a.ldq + h * kFD + d] * a.scale : 0.f;
  }
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f}, o0[2] = {0.f, 0.f}, o1[2] = {0.f, 0.f};
  const float* kg = a.k + static_cast<size_t>(b) * a.Sk * a.ldk + h * kFD;
  const float* vg = a.v + static_cast<size_t>(b) * a.Sk * a.ldv + h * kFD;
  for (int k0 = 0; k0 < a.Sk; k0 += kFChunk) {
    __syncthreads();  // previous chunk consumed (and qs written, first pass)
    for (int i = threadIdx.x; i < kFChunk * (kFD / 4); i += 256) {
      const int r = i >> 4, c = (i & 15) * 4;
      float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
      if (k0 + r < a.Sk) {
        kk = *reinterpret_cast<const float4*>(kg + static_cast<size_t>(k0 + r) * a.ldk + c);
        vv = *reinterpret_cast<const float4*>(vg + static_cast<size_t>(k0 + r) * a.ldv + c);
      }
      float* kd = Ks + r * kFKPitch + c;
      kd[0] = kk.x; kd[1] = kk.y; kd[2] = kk.z; kd[3] = kk.w;
      *reinterpret_cast<float4*>(Vs + r * kFD + c) = vv;
    }
    if (threadIdx.x < kFChunk) {
      const int j = k0 + threadIdx.x;
      float mk = -INFINITY;
      if (j < a.Sk) mk = (a.key_valid == nullptr || a.key_valid[static_cast<size_t>(b) * a.Sk + j]) ? 0.f : a.mask_value;
      kb[threadIdx.x] = mk;
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = warp * 2 + rr;
      const int q = q0 + r;
      if (q >= a.Sq) continue;  // warp-uniform
      const float* qr = qs + r * kFD;
      const float* pair = a.pair ? a.pair + (static_cast<size_t>(b) * a.Sq + q) * a.Sk : nullptr;
      float sc[2];
      float mx = -INFINITY;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int jl = t * 32 + lane;
        const float* kr = Ks + jl * kFKPitch;
        float acc = 0.f;
#pragma unroll 16
        for (int d = 0; d < kFD; ++d) acc = fmaf(qr[d], kr[d], acc);
        float s = acc + kb[jl];
        if (pair && k0 + jl < a.Sk) s += fmaf(pw, pair[k0 + jl], pb);
        sc[t] = s;
        mx = fmaxf(mx, s);
      }
      mx = warp_max(mx);
      const float m_new = fmaxf(m[rr], mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = __expf(m[rr] - m_use);  // exp(-inf) = 0 on the first chunk
      float psum = 0.f;
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const float e = __expf(sc[t] - m_use);
        ps[r * kFChunk + t * 32 + lane] = e;
        psum += e;
      }
      psum = warp_sum(psum);
      __syncwarp();
      float a0 = 0.f, a1 = 0.f;
      const float* pr = ps + r * kFChunk;
#pragma unroll 8
      for (int j = 0; j < kFChunk; ++j) {
        const float pj = pr[j];
        const float2 vv = *reinterpret_cast<const float2*>(Vs + j * kFD + lane * 2);
        a0 = fmaf(pj, vv.x, a0);
        a1 = fmaf(pj, vv.y, a1);
      }
      o0[rr] = fmaf(o0[rr], alpha, a0);
      o1[rr] = fmaf(o1[rr], alpha, a1);
      l[rr] = fmaf(l[rr], alpha, psum);
      m[rr] = m_new;
      __syncwarp();
    }
  }
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int q = q0 + warp * 2 + rr;
    if (q >= a.Sq) continue;
    const float inv = 1.0f / l[rr];
    *reinterpret_cast<float2*>(a.out + (static_cast<size_t>(b) * a.Sq + q) * a.ldo + h * kFD + lane * 2) =
        make_float2(o0[rr] * inv, o1[rr] * inv);
  }
}
}  // namespace

int attention_f32_fwd(int B, int heads, int Sq, int Sk, const float* q, int ldq, const float* k, int ldk, const float* v,
                      int ldv, float scale, const uint8_t* key_valid, float mask_value, const float* pair, float pair_w,
                      float pair_b, const float* pair_w_dev, const float* pair_b_dev, float* out, int ldo, cudaStream_t s) {
  ETP_REQUIRE(B > 0 && heads > 0 && Sq > 0 && Sk > 0, "attention_f32: empty problem");
  ETP_REQUIRE(q && k && v && out, "attention_f32: null argument");
  ETP_REQUIRE(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && ldo % 2 == 0, "attention_f32: pitches must keep rows 16-byte aligned");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  ETP_REQUIRE(al16(q) && al16(k) && al16(v) && al16(out), "attention_f32: pointers must be 16-byte aligned");
  ETP_REQUIRE(heads <= 65535 && B <= 65535, "attention_f32: grid limits");
  AttnF32 a;
  a.B = B; a.heads = heads; a.Sq = Sq; a.Sk = Sk; a.q = q; a.k = k; a.v = v; a.ldq = ldq; a.ldk = ldk; a.ldv = ldv;
  a.scale = scale; a.key_valid = key_valid; a.mask_value = mask_value; a.pair = pair; a.pair_w_dev = pair_w_dev;
  a.pair_b_dev = pair_b_dev; a.pair_w = pair_w; a.pair_b = pair_b; a.out = out; a.ldo = ldo;
  dim3 grid((Sq + kFRows - 1) / kFRows, heads, B);
  ETP_CHECK_CUDA(launch_pdl(attention_f32_kernel, grid, dim3(256), 0, s, a));
  ETP_LAUNCHED();
  return ETP_OK;
}

// ---------------------------------------------------------------------------------------------------
// step-level sequencing, all activations fp32
// ---------------------------------------------------------------------------------------------------
namespace {

struct Hp {
  bf16* a3 = nullptr;     // split A operand of the current GEMM, [rows, 3K]
  size_t a3_elems = 0;
  cudaStream_t s = nullptr;
};

// out_f32[rows, N] = act(A[rows, K] . W[N, K]^T + bias) (+ resid), W given as the hi|hi|lo image [N, 3K]
int linear_hp(const Hp& hp, const float* A, int rows, int K, const void* W3, int N, const float* bias, int act,
              const float* resid, float* out, int ld_out = 0) {
  ETP_REQUIRE(static_cast<size_t>(rows) * 3 * K <= hp.a3_elems, "high-precision forward: operand scratch too small");
  ETP_TRY(split3(A, hp.a3, rows, K, K, 0, hp.s));
  GemmArgs g;
  g.M = rows; g.N = N; g.K = 3 * K;
  g.A = hp.a3; g.lda = 3 * K;
  g.B = static_cast<const bf16*>(W3); g.ldb = 3 * K;
  g.bias = bias; g.act = act;
  g.resid = resid; g.ld_resid = N;
  g.out_f32 = out; g.ld_f32 = ld_out ? ld_out : N;
  return gemm(g, hp.s);
}

struct HpNavBufs {
  float *x = nullptr, *xa = nullptr, *xc = nullptr, *t = nullptr, *q = nullptr, *ctx = nullptr, *qkv = nullptr, *h = nullptr,
        *kv_all = nullptr, *stats = nullptr;
  Hp hp;
  void carve(Arena& ar, size_t rows, size_t kv_rows, int X) {
    x = ar.take<float>(rows * kH); xa = ar.take<float>(rows * kH); xc = ar.take<float>(rows * kH);
    t = ar.take<float>(rows * kH); q = ar.take<float>(rows * kH); ctx = ar.take<float>(rows * kH);
    qkv = ar.take<float>(rows * 3 * kH); h = ar.take<float>(rows * kI);
    kv_all = ar.take<float>(kv_rows * 2 * kH * (X > 0 ? X : 1));
    stats = ar.take<float>(rows * 2);
    const size_t a3 = 3 * ((rows * kI > kv_rows * kH) ? rows * kI : kv_rows * kH);
    hp.a3 = ar.take<bf16>(a3);
    hp.a3_elems = a3;
  }
};

// post-LN self-attention + FFN block on fp32 activations (BertAttention + BertIntermediate + BertOutput)
int self_ffn_block_hp(const etp_layer_weights& w, float eps, HpNavBufs& b, const float* a_in, int B, int S,
                      const uint8_t* key_valid, const float* pair, const float* pair_w, const float* pair_b, float* x_out) {
  const int rows = B * S;
  const Hp& hp = b.hp;
  ETP_TRY(linear_hp(hp, a_in, rows, kH, w.sqkv_w, 3 * kH, w.sqkv_b, 0, nullptr, b.qkv));
  ETP_TRY(attention_f32_fwd(B, kHeads, S, S, b.qkv, 3 * kH, b.qkv + kH, 3 * kH, b.qkv + 2 * kH, 3 * kH, 0.125f, key_valid,
                            -10000.0f, pair, 0.f, 0.f, pair_w, pair_b, b.ctx, kH, hp.s));
  ETP_TRY(linear_hp(hp, b.ctx, rows, kH, w.so_w, kH, w.so_b, 0, a_in, b.t));
  ETP_TRY(layernorm_fwd(b.t, w.sln_g, w.sln_b, eps, rows, kH, b.xc, nullptr, nullptr, nullptr, hp.s));
  ETP_TRY(linear_hp(hp, b.xc, rows, kH, w.f1_w, kI, w.f1_b, 1, nullptr, b.h));
  ETP_TRY(linear_hp(hp, b.h, rows, kI, w.f2_w, kH, w.f2_b, 0, b.xc, b.t));
  ETP_TRY(layernorm_fwd(b.t, w.fln_g, w.fln_b, eps, rows, kH, x_out, nullptr, nullptr, nullptr, hp.s));
  return ETP_OK;
}

}  // namespace

int forward_navigation_hp(const etp_nav_weights& w, const etp_nav_inputs& in, float* gmap_embeds, float* global_logits,
                          void* work, size_t work_bytes, cudaStream_t s) {
  const int B = in.B, N = in.N, L = in.L, X = w.num_x_layers;
  ETP_REQUIRE(B > 0 && N > 0 && L > 0 && X >= 0, "forward_navigation_hp: bad shape");
  ETP_REQUIRE(in.txt_embeds != nullptr, "forward_navigation_hp: fp32 txt_embeds required");
  Arena ar(work, work_bytes);
  HpNavBufs b;
  b.carve(ar, static_cast<size_t>(B) * N, static_cast<size_t>(B) * L, X);
  ETP_REQUIRE(ar.off <= work_bytes, "forward_navigation_hp: workspace too small");
  b.hp.s = s;
  const int rows = B * N;
  NodePackArgs np;
  np.rows = rows; np.img_fts = in.gmap_img_fts; np.step_ids = in.gmap_step_ids; np.pos_fts = in.gmap_pos_fts;
  np.pos_w = w.pos_w; np.pos_b = w.pos_b; np.pos_g = w.pos_g; np.pos_bb = w.pos_bb; np.step_emb = w.step_emb;
  np.x_f32 = X > 0 ? b.x : gmap_embeds;
  ETP_TRY(node_pack_fwd(np, s));
  if (X > 0)
    ETP_TRY(linear_hp(b.hp, in.txt_embeds, B * L, kH, w.xkv_all_w, X * 2 * kH, w.xkv_all_b, 0, nullptr, b.kv_all));
  const float* x = np.x_f32;
  const int ldkv = X * 2 * kH;
  for (int i = 0; i < X; ++i) {
    const etp_layer_weights& lw = w.layers[i];
    const float* kv = b.kv_all + static_cast<size_t>(i) * 2 * kH;
    ETP_TRY(linear_hp(b.hp, x, rows, kH, lw.xq_w, kH, lw.xq_b, 0, nullptr, b.q));
    ETP_TRY(attention_f32_fwd(B, kHeads, N, L, b.q, kH, kv, ldkv, kv + kH, ldkv, 0.125f, in.txt_masks, -10000.0f, nullptr, 0.f,
                              0.f, nullptr, nullptr, b.ctx, kH, s));
    ETP_TRY(linear_hp(b.hp, b.ctx, rows, kH, lw.xo_w, kH, lw.xo_b, 0, x, b.t));
    ETP_TRY(layernorm_fwd(b.t, lw.xln_g, lw.xln_b, w.ln_eps, rows, kH, b.xa, nullptr, nullptr, nullptr, s));
    float* x_out = (i == X - 1) ? gmap_embeds : b.x;
    ETP_TRY(self_ffn_block_hp(lw, w.ln_eps, b, b.xa, B, N, in.gmap_masks, w.sprel_w ? in.gmap_pair_dists : nullptr, w.sprel_w,
                              w.sprel_b, x_out));
    x = x_out;
  }
  ETP_TRY(linear_hp(b.hp, x, rows, kH, w.sap0_w, kH, w.sap0_b, 2, nullptr, b.t));
  ETP_TRY(sap_tail_fwd(b.t, w.sap_g, w.sap_bb, w.sap4_w, w.sap4_b, in.gmap_visited_masks, in.gmap_masks, rows, kH, global_logits,
                       b.stats, b.stats + rows, s));
  return ETP_OK;
}

static size_t hp_nav_bytes(int B, int N, int L, int X) {
  Arena ar(nullptr, ~size_t(0));
  HpNavBufs b;
  b.carve(ar, static_cast<size_t>(B) * N, static_cast<size_t>(B) * L, X);
  return ar.off;
}

int forward_txt_hp(const etp_txt_weights& w, const int64_t* txt_ids, const uint8_t* txt_masks, int B, int L,
                   float* txt_embeds, void* work, size_t work_bytes, cudaStream_t s) {
  const int NL = w.num_l_layers;
  ETP_REQUIRE(B > 0 && L > 0 && NL >= 0, "forward_txt_hp: bad shape");
  Arena ar(work, work_bytes);
  HpNavBufs b;
  b.carve(ar, static_cast<size_t>(B) * L, 0, 0);
  ETP_REQUIRE(ar.off <= work_bytes, "forward_txt_hp: workspace too small");
  b.hp.s = s;
  float* x = NL > 0 ? b.x : txt_embeds;
  ETP_TRY(embed_txt_fwd(txt_ids, w.word_emb, w.pos_emb, w.type_emb0, w.emb_g, w.emb_b, w.ln_eps, B, L, x, nullptr, nullptr,
                        b.stats, s));
  for (int i = 0; i < NL; ++i) {
    // ping-pong x <-> xa: the block reads its input as the residual of its first GEMM and writes x_out last
    float* x_out = (i == NL - 1) ? txt_embeds : (x == b.x ? b.xa : b.x);
    ETP_TRY(self_ffn_block_hp(w.layers[i], w.ln_eps, b, x, B, L, txt_masks, nullptr, nullptr, nullptr, x_out));
    x = x_out;
  }
  return ETP_OK;
}

namespace {
struct HpPanoBufs {
  float *rgb_lin, *dep_lin, *x, *x2, *y, *qkv, *ctx, *h, *stats;
  Hp hp;
  void carve(Arena& ar, size_t rows) {
    rgb_lin = ar.take<float>(rows * kH); dep_lin = ar.take<float>(rows * kH);
    x = ar.take<float>(rows * kH); x2 = ar.take<float>(rows * kH); y = ar.take<float>(rows * kH);
    qkv = ar.take<float>(rows * 3 * kH); ctx = ar.take<float>(rows * kH); h = ar.take<float>(rows * kI);
    stats = ar.take<float>(rows * 8);
    hp.a3 = ar.take<bf16>(rows * 3 * kI);
    hp.a3_elems = rows * 3 * kI;
  }
};
}  // namespace

int forward_panorama_hp(const etp_pano_weights& w, const etp_pano_inputs& in, float* pano_embeds, uint8_t* pano_masks,
                        void* work, size_t work_bytes, cudaStream_t s) {
  const int B = in.B, V = in.V, P = w.num_pano_layers;
  ETP_REQUIRE(B > 0 && V > 0 && P >= 0, "forward_panorama_hp: bad shape");
  Arena ar(work, work_bytes);
  HpPanoBufs b;
  b.carve(ar, static_cast<size_t>(B) * V);
  ETP_REQUIRE(ar.off <= work_bytes, "forward_panorama_hp: workspace too small");
  b.hp.s = s;
  const int rows = B * V;
  ETP_TRY(seq_mask(in.view_lens, B, V, pano_masks, s));
  ETP_TRY(linear_hp(b.hp, in.rgb_fts, rows, 512, w.img_w, kH, w.img_b, 0, nullptr, b.rgb_lin));
  if (w.dep_w) ETP_TRY(linear_hp(b.hp, in.dep_fts, rows, 128, w.dep_w, kH, w.dep_b, 0, nullptr, b.dep_lin));
  float* x = (P == 0) ? pano_embeds : b.x;
  PanoPackArgs pp;
  pp.rows = rows; pp.rgb_lin = b.rgb_lin; pp.dep_lin = w.dep_w ? b.dep_lin : nullptr; pp.loc_fts = in.loc_fts;
  pp.nav_types = in.nav_types; pp.loc_w = w.loc_w; pp.loc_b = w.loc_b;
  pp.img_g = w.img_g; pp.img_b = w.img_bb; pp.dep_g = w.dep_g; pp.dep_b = w.dep_bb; pp.loc_g = w.loc_g; pp.loc_bb = w.loc_bb;
  pp.out_g = w.out_g; pp.out_b = w.out_bb; pp.nav_emb = w.nav_emb; pp.tok_emb1 = w.tok_emb1;
  pp.x_f32 = x; pp.stats = b.stats;
  ETP_TRY(pano_pack_fwd(pp, s));
  for (int i = 0; i < P; ++i) {
    const etp_pano_layer_weights& lw = w.layers[i];
    float* x_mid = b.x2;
    ETP_TRY(layernorm_fwd(x, lw.n1_g, lw.n1_b, w.layer_eps, rows, kH, b.y, nullptr, nullptr, nullptr, s));
    ETP_TRY(linear_hp(b.hp, b.y, rows, kH, lw.in_w, 3 * kH, lw.in_b, 0, nullptr, b.qkv));
    ETP_TRY(attention_f32_fwd(B, kHeads, V, V, b.qkv, 3 * kH, b.qkv + kH, 3 * kH, b.qkv + 2 * kH, 3 * kH, 0.125f, pano_masks,
                              -INFINITY, nullptr, 0.f, 0.f, nullptr, nullptr, b.ctx, kH, s));
    ETP_TRY(linear_hp(b.hp, b.ctx, rows, kH, lw.out_w, kH, lw.out_b, 0, x, x_mid));
    ETP_TRY(layernorm_fwd(x_mid, lw.n2_g, lw.n2_b, w.layer_eps, rows, kH, b.y, nullptr, nullptr, nullptr, s));
    ETP_TRY(linear_hp(b.hp, b.y, rows, kH, lw.l1_w, kI, lw.l1_b, 1, nullptr, b.h));
    ETP_TRY(linear_hp(b.hp, b.h, rows, kI, lw.l2_w, kH, lw.l2_b, 0, x_mid, b.x));
    x = b.x;
  }
  if (P > 0) ETP_TRY(layernorm_fwd(x, w.fin_g, w.fin_b, 1e-12f, rows, kH, pano_embeds, nullptr, nullptr, nullptr, s));
  return ETP_OK;
}

}  // namespace etp

using namespace etp;
#define ETP_API __attribute__((visibility("default")))
static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

ETP_API int etp_split3(const float* x, void* y, int64_t rows, int32_t K, int32_t form, void* stream) {
  return split3(x, static_cast<bf16*>(y), rows, K, K, form, S(stream));
}

ETP_API int etp_attention_f32_fwd(const etp_attn_args* g, void* stream) {
  ETP_REQUIRE(g != nullptr, "etp_attention_f32_fwd: null args");
  return attention_f32_fwd(g->B, g->heads, g->Sq, g->Sk, static_cast<const float*>(g->q), g->ldq,
                           static_cast<const float*>(g->k), g->ldk, static_cast<const float*>(g->v), g->ldv, g->scale,
                           g->key_valid, g->mask_value, g->pair, g->pair_w, g->pair_b, g->pair_w_dev, g->pair_b_dev,
                           static_cast<float*>(g->out), g->ldo, S(stream));
}

ETP_API size_t etp_hp_nav_work_bytes(int32_t B, int32_t N, int32_t L, int32_t X) { return hp_nav_bytes(B, N, L, X); }
ETP_API size_t etp_hp_txt_work_bytes(int32_t B, int32_t L) { return hp_nav_bytes(B, L, 0, 0); }
ETP_API size_t etp_hp_pano_work_bytes(int32_t B, int32_t V) {
  Arena ar(nullptr, ~size_t(0));
  HpPanoBufs b;
  b.carve(ar, static_cast<size_t>(B) * V);
  return ar.off;
}

ETP_API int etp_forward_navigation_hp(const etp_nav_weights* w, const etp_nav_inputs* in, float* gmap_embeds,
                                      float* global_logits, void* work, size_t work_bytes, void* stream) {
  ETP_REQUIRE(w && in && gmap_embeds && global_logits && work, "etp_forward_navigation_hp: null argument");
  return forward_navigation_hp(*w, *in, gmap_embeds, global_logits, work, work_bytes, S(stream));
}
ETP_API int etp_forward_panorama_hp(const etp_pano_weights* w, const etp_pano_inputs* in, float* pano_embeds,
                                    uint8_t* pano_masks, void* work, size_t work_bytes, void* stream) {
  ETP_REQUIRE(w && in && pano_embeds && pano_masks && work, "etp_forward_panorama_hp: null argument");
  return forward_panorama_hp(*w, *in, pano_embeds, pano_masks, work, work_bytes, S(stream));
}
ETP_API int etp_forward_txt_hp(const etp_txt_weights* w, const int64_t* txt_ids, const uint8_t* txt_masks, int32_t B,
                               int32_t L, float* txt_embeds, void* work, size_t work_bytes, void* stream) {
  ETP_REQUIRE(w && txt_ids && txt_masks && txt_embeds && work, "etp_forward_txt_hp: null argument");
  return forward_txt_hp(*w, txt_ids, txt_masks, B, L, txt_embeds, work, work_bytes, S(stream));
}

}  // extern "C"
