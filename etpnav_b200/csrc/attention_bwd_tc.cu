// Attention backward (head dim 64) on tcgen05 tensor cores, for query sequences that fit one 128-row tile
// (graph nodes: <= 128).  PERSISTENT: one CTA per SM walks over (batch, head) items; keys in blocks of 128.
//
//   per key block ("step"):  S = Q.K^T, dP = dO.V^T                 (2 x [M=128,N=128,K=64] into TMEM)
//     8 softmax-backward warps — thread = (query row, 64-key half): P = exp2(s2 - lse2), dS = P * (dP - D) (x scale),
//     both written to shared memory as bf16 in the 128B-swizzled [query, key] layout, which the tensor core
//     reads K-major (dQ += dS.K) AND MN-major (dV = P^T.dO, dK = dS^T.Q) — the same bytes, two descriptors;
//     sprel_linear gradients (sum dS*pair, sum dS) accumulate per thread over the CTA's whole item list;
//   then  dV_blk = P^T.dO, dK_blk = dS^T.Q  ([M=128 keys, N=64, K=128 queries]),  dQ += dS.K  (accumulating
//   in TMEM across key blocks); dK/dV rows are read back (thread = key; one warp group takes dV, the other dK)
//   and stored as bf16; dQ after the last block.  Q and dO tiles double as MN-major B operands, K as both
//   K-major (S) and MN-major (dQ) B.
//
// The control thread runs one step ahead of the math: Q/dO and K/V tiles are double-buffered, so the TMA
// loads of step s+1 (possibly the next item) are in flight while the warps work on step s, and S/dP of
// step s+1 are issued right behind the dV/dK/dQ MMAs of step s.  TMEM allocation, barrier set-up and the
// launch are paid once per CTA instead of once per (batch, head).
// Gradient counterpart of BertOutAttention / BertSelfAttention (vilmodel_cmt.py:325-352, 103-141, 391-393).
#include "common.cuh"
#include "host.h"
#include "ops.h"

namespace etp {

namespace {

constexpr int kBQ = 128;
constexpr int kBK = 128;
constexpr int kD = 64;
constexpr int kTile = kBQ * kD * 2;     // 16 KB: Q, dO, K, V tiles
constexpr int kPBytes = kBQ * kBK * 2;  // 32 KB: P and dS
constexpr int kWarps = 8;               // softmax-backward warps
constexpr int kMathThreads = kWarps * 32;
constexpr int kThreads = kMathThreads + 32;  // + control warp
constexpr int kSmemBytes = 8 * kTile + 2 * kPBytes + 2 * kBK * 4 + 64 + 1024 + 256;
constexpr uint32_t kTmemCols = 512;  // S [0,128) dP [128,256) dV [256,320) dK [320,384) dQ [384,448)
constexpr float kLog2e = 1.4426950408889634f;

struct BwdDev {
  int B, heads, Sq, Sk;
  float scale;
  const uint8_t* key_valid;
  float mask_value;
  const float* pair;
  float pair_w, pair_b;
  const float* pair_w_dev;
  const float* pair_b_dev;
  const bf16* out; int ldo;
  const bf16* dout; int lddo;
  const float* lse;
  bf16 *dq, *dk, *dv;
  int lddq, lddk, lddv;
  float *dpair_w, *dpair_b;
};

ETP_DEVICE uint4 pack8(const float* f) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
ETP_DEVICE uint4 pack8u(const uint32_t* u) {
  return make_uint4(pack_bf16x2(__uint_as_float(u[0]), __uint_as_float(u[1])),
                    pack_bf16x2(__uint_as_float(u[2]), __uint_as_float(u[3])),
                    pack_bf16x2(__uint_as_float(u[4]), __uint_as_float(u[5])),
                    pack_bf16x2(__uint_as_float(u[6]), __uint_as_float(u[7])));
}

__global__ void __launch_bounds__(kThreads, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                        const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const BwdDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                 // [2][16 KB]
  uint8_t* sDO = sQ + 2 * kTile;      // [2][16 KB]
  uint8_t* sK = sDO + 2 * kTile;      // [2][16 KB]
  uint8_t* sV = sK + 2 * kTile;       // [2][16 KB]
  uint8_t* sP = sV + 2 * kTile;
  uint8_t* sDS = sP + kPBytes;
  float* sKb = reinterpret_cast<float*>(sDS + kPBytes);  // [2][128] per-key additive mask (log2 domain)
  float* sRed = sKb + 2 * kBK;                            // [16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + 16);
  uint64_t* q_full = bars + 0;    // [2]
  uint64_t* kv_full = bars + 2;   // [2]
  uint64_t* sdp_ready = bars + 4;
  uint64_t* pds_ready = bars + 5;
  uint64_t* g_ready = bars + 6;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = (p.Sk + kBK - 1) / kBK;
  const int n_items = p.B * p.heads;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(&q_full[0], 1); mbar_init(&q_full[1], 1);
    mbar_init(&kv_full[0], 1); mbar_init(&kv_full[1], 1);
    mbar_init(sdp_ready, 1);
    mbar_init(pds_ready, kMathThreads);
    mbar_init(g_ready, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  griddep_launch();  // PDL: the next kernel may start its own prologue
  griddep_wait();    // previous kernel complete; nothing above touched global memory or TMEM
  if (warp == kWarps) {
    tmem_alloc(tmem_ptr, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *tmem_ptr;
  const uint32_t tS = tm, tDP = tm + 128, tDV = tm + 256, tDK = tm + 320, tDQ = tm + 384;

  if (warp == kWarps) {
    // ======================= control warp: TMA + MMA issue (one lane), one step ahead =======================
    if (lane == 0 && static_cast<int>(blockIdx.x) < n_items) {
      constexpr uint32_t id_s = make_idesc_bf16(128, 128, 0, 0);   // [q, d] x [key, d]^T
      constexpr uint32_t id_kv = make_idesc_bf16(128, 64, 1, 1);   // A = P/dS read MN-major, B = dO/Q MN-major
      constexpr uint32_t id_q = make_idesc_bf16(128, 64, 0, 1);    // A = dS K-major, B = K MN-major
      const uint32_t aP = smem_u32(sP), aDS = smem_u32(sDS);
      auto load_q = [&](int item, int st) {
        const int b = item / p.heads, h = item % p.heads;
        mbar_arrive_expect_tx(&q_full[st], 2 * kTile);
        tma_load_3d(sQ + st * kTile, &tmQ, &q_full[st], h * kD, 0, b);
        tma_load_3d(sDO + st * kTile, &tmDO, &q_full[st], h * kD, 0, b);
      };
      auto load_kv = [&](int item, int j, int st) {
        const int b = item / p.heads, h = item % p.heads;
        mbar_arrive_expect_tx(&kv_full[st], 2 * kTile);
        tma_load_3d(sK + st * kTile, &tmK, &kv_full[st], h * kD, j * kBK, b);
        tma_load_3d(sV + st * kTile, &tmV, &kv_full[st], h * kD, j * kBK, b);
      };
      load_q(blockIdx.x, 0);
      load_kv(blockIdx.x, 0, 0);
      int s = 0;   // step counter (item, key block) of this CTA
      int ii = 0;  // item counter of this CTA
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ii) {
        const int qst = ii & 1;
        const uint32_t aQ = smem_u32(sQ + qst * kTile), aDO = smem_u32(sDO + qst * kTile);
        for (int j = 0; j < nblk; ++j, ++s) {
          const int kst = s & 1;
          const uint32_t aK = smem_u32(sK + kst * kTile), aV = smem_u32(sV + kst * kTile);
          if (j == 0) mbar_wait(&q_full[qst], (ii >> 1) & 1);
          mbar_wait(&kv_full[kst], (s >> 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < kD / 16; ++k)
            umma_bf16(tS, make_smem_desc(aQ + k * 32, 16, 1024), make_smem_desc(aK + k * 32, 16, 1024), id_s, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < kD / 16; ++k)
            umma_bf16(tDP, make_smem_desc(aDO + k * 32, 16, 1024), make_smem_desc(aV + k * 32, 16, 1024), id_s, k > 0 ? 1u : 0u);
          umma_commit(sdp_ready);
          // prefetch the tiles of the next step; its buffers were last read by the MMAs of step s-1
          const bool more_blocks = j + 1 < nblk;
          const int next_item = item + gridDim.x;
          if (more_blocks || next_item < n_items) {
            if (s >= 1) mbar_wait(g_ready, (s - 1) & 1);
            if (more_blocks) {
              load_kv(item, j + 1, (s + 1) & 1);
            } else {
              load_q(next_item, (ii + 1) & 1);
              load_kv(next_item, 0, (s + 1) & 1);
            }
          }
          mbar_wait(pds_ready, s & 1);
          tc_fence_after();
          // reduction over the 128 queries, 16 per MMA: A panels are [query rows of 128 B] -> advance 16 rows = 2048 B,
          // two 64-key panels 16 KB apart (LBO), 8-row groups 1 KB apart (SBO); B = dO / Q tile read MN-major.
#pragma unroll
          for (int k = 0; k < kBQ / 16; ++k) {
            umma_bf16(tDV, make_smem_desc(aP + k * 2048, 16384, 1024), make_smem_desc(aDO + k * 2048, 8192, 1024), id_kv,
                      k > 0 ? 1u : 0u);
            umma_bf16(tDK, make_smem_desc(aDS + k * 2048, 16384, 1024), make_smem_desc(aQ + k * 2048, 8192, 1024), id_kv,
                      k > 0 ? 1u : 0u);
          }
          // dQ += dS.K : reduction over the 128 keys of this block
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_bf16(tDQ, make_smem_desc(aDS + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                      make_smem_desc(aK + k * 2048, 8192, 1024), id_q, (j > 0 || k > 0) ? 1u : 0u);
          umma_commit(g_ready);
        }
      }
    }
  } else {
    // ======================= softmax-backward warps =======================
    const int quad = warp & 3;            // TMEM lane quadrant
    const int hh = warp >> 2;             // which 64 keys of the block (softmax phase) / dV vs dK (read-back phase)
    const int r = quad * 32 + lane;       // query row (softmax phase) / key row of the block (dK, dV read-back)
    const bool qv = r < p.Sq;
    const bool warp_live = quad * 32 < p.Sq;  // some row of this warp is a real query
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const float pw_raw = p.pair_w_dev ? __ldg(p.pair_w_dev) : p.pair_w;
    const float pb_raw = p.pair_b_dev ? __ldg(p.pair_b_dev) : p.pair_b;
    const float pw = pw_raw * kLog2e, pb = pb_raw * kLog2e;
    const float sl2 = p.scale * kLog2e;
    const float mask2 = p.mask_value * kLog2e;
    const bool pair_vec = (p.Sk & 3) == 0;  // rows of the pair bias are 16-byte aligned
    float wsum = 0.f, bsum = 0.f;
    int s = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int b = item / p.heads, h = item % p.heads;
      const uint8_t* kvalid = p.key_valid ? p.key_valid + static_cast<size_t>(b) * p.Sk : nullptr;
      const float* pair_row = p.pair ? p.pair + (static_cast<size_t>(b) * p.Sq + r) * p.Sk : nullptr;
      // D = sum_d dO * O and the saved log-sum-exp of this query row
      float Dv = 0.f, lse2 = 0.f;
      if (qv) {
        const size_t row = static_cast<size_t>(b) * p.Sq + r;
        const uint4* po = reinterpret_cast<const uint4*>(p.out + row * p.ldo + h * kD);
        const uint4* pd = reinterpret_cast<const uint4*>(p.dout + row * p.lddo + h * kD);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 a = __ldg(po + i), c = __ldg(pd + i);
          const __nv_bfloat162* ha = reinterpret_cast<const __nv_bfloat162*>(&a);
          const __nv_bfloat162* hc = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 fa = __bfloat1622float2(ha[t]), fc = __bfloat1622float2(hc[t]);
            Dv += fa.x * fc.x + fa.y * fc.y;
          }
        }
        lse2 = p.lse[(static_cast<size_t>(b) * p.heads + h) * p.Sq + r] * kLog2e;
      }

      for (int j = 0; j < nblk; ++j, ++s) {
        const uint32_t ph = s & 1;
        const int k0 = j * kBK;
        float* kb = sKb + (s & 1) * kBK;
        if (threadIdx.x < kBK) {
          const int k = k0 + static_cast<int>(threadIdx.x);
          float v = -INFINITY;
          if (k < p.Sk) v = (kvalid && !kvalid[k]) ? mask2 : 0.f;
          kb[threadIdx.x] = v;
        }
        // pair bias of this thread's row for its first 32 keys: in flight while the S / dP MMAs run
        float pv[32];
        auto load_pair = [&](int c0) {
          const int key0 = k0 + c0;
          if (pair_row != nullptr && qv) {
            if (pair_vec) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
                if (key0 + i < p.Sk) f = __ldg(reinterpret_cast<const float4*>(pair_row + key0 + i));
                pv[i] = f.x; pv[i + 1] = f.y; pv[i + 2] = f.z; pv[i + 3] = f.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) pv[i] = (key0 + i < p.Sk) ? __ldg(pair_row + key0 + i) : 0.f;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) pv[i] = 0.f;
          }
        };
        load_pair(hh * 64);
        named_bar_sync(1, kMathThreads);  // key mask visible
        mbar_wait(sdp_ready, ph);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int c = hh * 64 + cc * 32;  // first key column of this chunk inside the block
          uint32_t vs[32], vd[32];
          float pe[32], de[32];
          if (warp_live) {  // warp-uniform: tcgen05.ld is a warp-collective instruction
            tmem_ld32(tS + lane_sel + c, vs);
            tmem_ld32(tDP + lane_sel + c, vd);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float sc = fmaf(__uint_as_float(vs[i]), sl2, kb[c + i]);
              if (pair_row) sc += fmaf(pw, pv[i], pb);
              const float pr = qv ? exp2f(sc - lse2) : 0.f;
              const float ds = pr * (__uint_as_float(vd[i]) - Dv);
              wsum = fmaf(ds, pv[i], wsum);
              bsum += ds;
              pe[i] = pr;
              de[i] = ds * p.scale;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) { pe[i] = 0.f; de[i] = 0.f; }
          }
          if (cc == 0) load_pair(hh * 64 + 32);  // next chunk's bias while this one is packed and stored
          uint8_t* prow = sP + (c >> 6) * 16384 + r * 128;
          uint8_t* drow = sDS + (c >> 6) * 16384 + r * 128;
          const int ch0 = (c & 63) >> 3;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int off = (((ch0 + g) ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(prow + off) = pack8(pe + 8 * g);
            *reinterpret_cast<uint4*>(drow + off) = pack8(de + 8 * g);
          }
        }
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(pds_ready);
        // dK / dV of this key block: thread r now owns key k0 + r; warps 0-3 store dV, warps 4-7 store dK
        mbar_wait(g_ready, ph);
        tc_fence_after();
        const int key = k0 + r;
        {
          const uint32_t tsrc = (hh == 0 ? tDV : tDK) + lane_sel;
          bf16* gbase = (hh == 0) ? p.dv : p.dk;
          const int ldg = (hh == 0) ? p.lddv : p.lddk;
#pragma unroll
          for (int c = 0; c < kD; c += 32) {
            uint32_t vv[32];
            tmem_ld32(tsrc + c, vv);
            tmem_ld_wait();
            if (key < p.Sk) {
              bf16* g = gbase + (static_cast<size_t>(b) * p.Sk + key) * ldg + h * kD + c;
#pragma unroll
              for (int i = 0; i < 32; i += 8) *reinterpret_cast<uint4*>(g + i) = pack8u(vv + i);
            }
          }
        }
        if (j == nblk - 1) {
          // dQ of this item (complete with this step's commit): each warp group stores 32 of the 64 columns
          uint32_t v[32];
          tmem_ld32(tDQ + lane_sel + hh * 32, v);
          tmem_ld_wait();
          if (qv) {
            bf16* gq = p.dq + (static_cast<size_t>(b) * p.Sq + r) * p.lddq + h * kD + hh * 32;
#pragma unroll
            for (int i = 0; i < 32; i += 8) *reinterpret_cast<uint4*>(gq + i) = pack8u(v + i);
          }
        }
        tc_fence_before();
      }
    }
    if (p.dpair_w) {
      wsum = warp_sum(wsum);
      bsum = warp_sum(bsum);
      if (lane == 0) { sRed[warp] = wsum; sRed[8 + warp] = bsum; }
      named_bar_sync(1, kMathThreads);
      if (threadIdx.x == 0) {
        float a = 0.f, c = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) { a += sRed[w]; c += sRed[8 + w]; }
        atomicAdd(p.dpair_w, a);
        atomicAdd(p.dpair_b, c);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarps) {
    tc_fence_after();
    tmem_dealloc(tm, kTmemCols);
  }
}

}  // namespace

bool attention_bwd_tc_supported(const AttnBwdArgs& a) {
  if (a.Sq > kBQ) return false;
  if (a.Sq < 32 && a.Sk < 32) return false;  // tiny problems: CUDA-core kernels
  if (a.ldq % 8 || a.ldk % 8 || a.ldv % 8 || a.lddo % 8 || a.ldo % 8 || a.lddq % 8 || a.lddk % 8 || a.lddv % 8) return false;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return al16(a.q) && al16(a.k) && al16(a.v) && al16(a.out) && al16(a.dout) && al16(a.dq) && al16(a.dk) && al16(a.dv) &&
         (a.pair == nullptr || al16(a.pair));
}

int attention_bwd_tc(const AttnBwdArgs& a, cudaStream_t stream) {
  ETP_REQUIRE(attention_bwd_tc_supported(a), "attention_bwd_tc: unsupported shape / layout");
  CUtensorMap tq, tdo, tk, tv;
  const uint64_t W = static_cast<uint64_t>(a.heads) * kD;
  int rc = get_tmap_3d(a.q, W, a.Sq, a.B, a.ldq, static_cast<uint64_t>(a.Sq) * a.ldq, kD, kBQ, &tq);
  if (rc) return rc;
  rc = get_tmap_3d(a.dout, W, a.Sq, a.B, a.lddo, static_cast<uint64_t>(a.Sq) * a.lddo, kD, kBQ, &tdo);
  if (rc) return rc;
  rc = get_tmap_3d(a.k, W, a.Sk, a.B, a.ldk, static_cast<uint64_t>(a.Sk) * a.ldk, kD, kBK, &tk);
  if (rc) return rc;
  rc = get_tmap_3d(a.v, W, a.Sk, a.B, a.ldv, static_cast<uint64_t>(a.Sk) * a.ldv, kD, kBK, &tv);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    ETP_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  BwdDev d;
  d.B = a.B; d.heads = a.heads; d.Sq = a.Sq; d.Sk = a.Sk; d.scale = a.scale; d.key_valid = a.key_valid;
  d.mask_value = a.mask_value; d.pair = a.pair; d.pair_w = a.pair_w; d.pair_b = a.pair_b; d.pair_w_dev = a.pair_w_dev;
  d.pair_b_dev = a.pair_b_dev; d.out = a.out; d.ldo = a.ldo; d.dout = a.dout; d.lddo = a.lddo; d.lse = a.lse;
  d.dq = a.dq; d.dk = a.dk; d.dv = a.dv; d.lddq = a.lddq; d.lddk = a.lddk; d.lddv = a.lddv;
  d.dpair_w = a.dpair_w; d.dpair_b = a.dpair_b;
  const int items = a.B * a.heads;
  const int grid = items < num_sms() ? items : num_sms();
  ETP_CHECK_CUDA(launch_pdl(attention_bwd_tc_kernel, dim3(grid), dim3(kThreads), kSmemBytes, stream, tq, tdo, tk, tv, d));
  ETP_LAUNCHED();
  return ETP_OK;
}

}  // namespace etp
