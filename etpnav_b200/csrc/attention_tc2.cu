// Multi-head attention forward (head dim 64) on tcgen05 tensor cores — second generation: TWO CTAs PER SM.
//
// attention_tc.cu keeps ONE (batch, head, query-tile) item in flight per SM: 16 softmax warps that all sit in the same
// phase (wait for S — read TMEM — exchange the row maximum — exp / dropout — write P — wait for P.V — read O), so the SM
// issues instructions 37-45 % of the time (profiles/r02_attention_fwd_ncu.txt: long-scoreboard, barrier and wait stalls
// while every warp waits for the same tensor-core / TMEM round trip) although the tensor pipe is 3-7 % busy: the kernel
// is bound by that per-item latency chain, not by any unit.  Here a CTA is half the size — 8 softmax warps (thread =
// query row x 64-key half of the 128-key step) + one control warp, 100 KB of shared memory (Q single, K double-, V
// single-buffered, P), 256 TMEM columns (S 128 + O 64) — so that two CTAs are resident per SM and one CTA's latency
// chain is covered by the other's arithmetic.  Pipelining INSIDE a CTA is reduced to what is free: K of step s+1 and
// Q of the next item are prefetched, S(s+1) is issued right behind P.V(s).
//
// Two resident 9-warp CTAs leave 96 registers per thread, so nothing row-sized lives in registers across a step: the softmax reads S twice from TMEM — pass A takes the row maximum, pass B
// recomputes the biased scores and writes P = exp2(s - m) (bf16, the 128B-swizzled K-major layout the P.V MMA reads) —
// and the output accumulates IN TMEM across key blocks (P.V with accumulate = 1); when a later block raises the row
// maximum the 32 columns a thread owns are rescaled in place (tcgen05.ld -> x alpha -> tcgen05.st) before that block's
// P.V is issued.  One-block items (self-attention over <= 128 nodes, instructions of <= 128 tokens) never rescale.
// Masks, pair bias in the log2 domain, dropout element indices, the TMA-stored output tile and lse follow
// attention_tc.cu, which stays as the ETP_ATTN_V2=0 variant (A/B measurements, tests run both).
// Replaces BertOutAttention / BertSelfAttention's softmax(QK^T/8 + mask [+ sprel]) V (vilmodel_cmt.py:325-352, 103-141).
#include <cstdlib>

#include "common.cuh"
#include "host.h"
#include "ops.h"

namespace etp {

namespace {

constexpr int kBQ = 128;
constexpr int kBK = 128;
constexpr int kD = 64;
constexpr int kTile = kBQ * kD * 2;     // 16 KB
constexpr int kPBytes = kBQ * kBK * 2;  // 32 KB
constexpr int kMathWarps = 8;
constexpr int kMathThreads = kMathWarps * 32;
constexpr int kThreads = kMathThreads + 32;
constexpr int kSmemBytes = 4 * kTile + kPBytes + 2 * kBK * 4 + 2 * kBQ * 4 + 2 * kBQ * 4 + 256 + 1024;
constexpr uint32_t kTmemCols = 256;  // S [0,128)  O [128,192)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct Attn2Dev {
  int B, heads, Sq, Sk;
  float scale;
  const uint8_t* key_valid;
  float mask_value;
  const float* pair;
  float pair_w, pair_b;
  const float* pair_w_dev;
  const float* pair_b_dev;
  float* lse;
  Drop drop;
  const int32_t* kv_rows;
};

ETP_DEVICE float ex2a(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
ETP_DEVICE uint4 pack8f(const float* f) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

template <bool kPair>
__global__ void __launch_bounds__(kThreads, 2)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const Attn2Dev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;              // 16 KB
  uint8_t* sK = sQ + kTile;        // [2][16 KB]
  uint8_t* sV = sK + 2 * kTile;    // 16 KB
  uint8_t* sP = sV + kTile;        // 32 KB; its first 16 KB double as the output tile of the TMA store
  float* sKb = reinterpret_cast<float*>(sP + kPBytes);  // [2][128] per-key additive bias (log2 domain)
  float* sMax = sKb + 2 * kBK;                           // [2][128] half-row maxima of the current step
  float* sL = sMax + 2 * kBQ;                            // [2][128] half-row sums at item end
  uint64_t* bars = reinterpret_cast<uint64_t*>(sL + 2 * kBQ);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* v_full = bars + 3;
  uint64_t* s_ready = bars + 4;
  uint64_t* p_ready = bars + 5;
  uint64_t* o_ready = bars + 6;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = (p.Sk + kBK - 1) / kBK;
  const int qtiles = (p.Sq + kBQ - 1) / kBQ;
  const int n_items = p.B * p.heads * qtiles;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    mbar_init(q_full, 1);
    mbar_init(&k_full[0], 1); mbar_init(&k_full[1], 1);
    mbar_init(v_full, 1);
    mbar_init(s_ready, 1);
    mbar_init(p_ready, kMathThreads);
    mbar_init(o_ready, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  griddep_launch();  // PDL: the next kernel may start its own prologue
  griddep_wait();    // previous kernel complete; nothing above touched global memory or TMEM
  if (warp == kMathWarps) {
    tmem_alloc(tmem_ptr, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tS = *tmem_ptr;
  const uint32_t tO = tS + 128;

  auto item_bhq = [&](int item, int& b, int& h, int& q0) {
    const int qt = item % qtiles;
    const int bh = item / qtiles;
    b = bh / p.heads; h = bh % p.heads; q0 = qt * kBQ;
  };

  if (warp == kMathWarps) {
    // ======================= control warp: TMA + MMA issue (one lane) =======================
    if (lane == 0 && static_cast<int>(blockIdx.x) < n_items) {
      constexpr uint32_t idesc_s = make_idesc_bf16(kBQ, kBK, 0, 0);  // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = make_idesc_bf16(kBQ, kD, 0, 1);   // P (K-major) x V (MN-major)
      const uint64_t dp0 = make_smem_desc(smem_u32(sP), 16, 1024);
      const uint64_t dq0 = make_smem_desc(smem_u32(sQ), 16, 1024);
      const uint64_t dv0 = make_smem_desc(smem_u32(sV), 8192, 1024);
      auto load_q = [&](int item) {
        int b, h, q0;
        item_bhq(item, b, h, q0);
        mbar_arrive_expect_tx(q_full, kTile);
        tma_load_3d(sQ, &tmQ, q_full, h * kD, q0, b);
      };
      auto load_k = [&](int item, int j, int st) {
        int b, h, q0;
        item_bhq(item, b, h, q0);
        const int kvb = p.kv_rows ? __ldg(p.kv_rows + b) : b;
        mbar_arrive_expect_tx(&k_full[st], kTile);
        tma_load_3d(sK + st * kTile, &tmK, &k_full[st], h * kD, j * kBK, kvb);
      };
      auto load_v = [&](int item, int j) {
        int b, h, q0;
        item_bhq(item, b, h, q0);
        const int kvb = p.kv_rows ? __ldg(p.kv_rows + b) : b;
        mbar_arrive_expect_tx(v_full, kTile);
        tma_load_3d(sV, &tmV, v_full, h * kD, j * kBK, kvb);
      };
      // S of step s (item counter ii): waits for its tiles, four MMAs.  The S columns are free: the softmax warps
      // finished reading S(s-1) before they arrived on p_ready(s-1), which this thread has observed.
      auto issue_s = [&](int ii, int s, bool first_block) {
        if (first_block) mbar_wait(q_full, ii & 1);
        mbar_wait(&k_full[s & 1], (s >> 1) & 1);
        tc_fence_after();
        const uint64_t dk0 = make_smem_desc(smem_u32(sK + (s & 1) * kTile), 16, 1024);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          umma_bf16_lh(tS, desc_lo(dq0) + 2 * k, desc_hi(dq0), desc_lo(dk0) + 2 * k, desc_hi(dk0), idesc_s, k > 0 ? 1u : 0u);
        umma_commit(s_ready);
      };
      load_q(blockIdx.x);
      load_k(blockIdx.x, 0, 0);
      load_v(blockIdx.x, 0);
      issue_s(0, 0, true);
      int s = 0, ii = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ii) {
        for (int j = 0; j < nblk; ++j, ++s) {
          const bool more_blocks = j + 1 < nblk;
          const int next_item = item + gridDim.x;
          const bool has_next = more_blocks || next_item < n_items;
          if (has_next) {
            // K of step s+1: its stage was read by S(s-1), long complete
            if (more_blocks) load_k(item, j + 1, (s + 1) & 1);
            else             load_k(next_item, 0, (s + 1) & 1);
            if (!more_blocks) {
              // Q of the next item: the Q tile is free once S of this item's last block has been computed
              mbar_wait(s_ready, s & 1);
              load_q(next_item);
            }
          }
          // O_blk = P.V of this step
          mbar_wait(p_ready, s & 1);
          mbar_wait(v_full, s & 1);
          tc_fence_after();
          const int kgroups = (min(kBK, p.Sk - j * kBK) + 15) >> 4;  // P is 0 beyond the real keys
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            if (k < kgroups)
              umma_bf16_lh(tO, desc_lo(dp0) + (k >> 2) * 1024 + (k & 3) * 2, desc_hi(dp0), desc_lo(dv0) + 128 * k, desc_hi(dv0),
                           idesc_o, (j > 0 || k > 0) ? 1u : 0u);   // O accumulates in TMEM over the item's key blocks
          umma_commit(o_ready);
          if (has_next) {
            issue_s(more_blocks ? ii : ii + 1, s + 1, !more_blocks);  // runs behind P.V(s)
            mbar_wait(o_ready, s & 1);                                // V tile free
            if (more_blocks) load_v(item, j + 1);
            else             load_v(next_item, 0);
          }
        }
      }
    }
  } else {
    // ======================= softmax warps: thread = (query row, 64-key half of the step) =======================
    const int quad = warp & 3;       // TMEM lane quadrant
    const int wq = warp >> 2;        // which 64 keys of the step / which 32 output columns
    const int r = quad * 32 + lane;  // row in the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const float pw2 = (p.pair_w_dev ? __ldg(p.pair_w_dev) : p.pair_w) * kLog2e;
    const float pb2 = (p.pair_b_dev ? __ldg(p.pair_b_dev) : p.pair_b) * kLog2e;
    const float sl2 = p.scale * kLog2e;
    const float mask2 = p.mask_value * kLog2e;
    const bool pair_vec = (p.Sk & 3) == 0;  // rows of the pair bias are 16-byte aligned

    // biased, log2-domain scores of 32 keys [key0, key0 + 32) of this thread's row from the raw TMEM values
    auto bias32 = [&](uint32_t (&vs)[32], const float* kb32, const float* pair_row, int key0) {
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 kb4 = *reinterpret_cast<const float4*>(kb32 + i);
        float pv[4] = {0.f, 0.f, 0.f, 0.f};
        if constexpr (kPair) {
          if (pair_row != nullptr) {
            if (pair_vec) {
              if (key0 + i < p.Sk) {
                const float4 f = __ldg(reinterpret_cast<const float4*>(pair_row + key0 + i));
                pv[0] = f.x; pv[1] = f.y; pv[2] = f.z; pv[3] = f.w;
              }
            } else {
#pragma unroll
              for (int t = 0; t < 4; ++t) pv[t] = (key0 + i + t < p.Sk) ? __ldg(pair_row + key0 + i + t) : 0.f;
            }
          }
        }
        const float kbv[4] = {kb4.x, kb4.y, kb4.z, kb4.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          float x = fmaf(__uint_as_float(vs[i + t]), sl2, kbv[t]);
          if constexpr (kPair) x = fmaf(pw2, pv[t], x);
          vs[i + t] = __float_as_uint(x);
        }
      }
    };

    int s = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int b, h, q0;
      item_bhq(item, b, h, q0);
      const int q = q0 + r;
      // a warp whose 32 rows all lie past the last query (rows 96..127 of an 80-node map) has nothing to compute: its P
      // rows only feed output rows the TMA store clips.  It keeps every barrier / phase, skips TMEM reads and the math.
      const bool warp_live = q0 + quad * 32 < p.Sq;  // warp-uniform: tcgen05.ld / st are warp collectives
      const float* pair_row = (kPair && q < p.Sq) ? p.pair + (static_cast<size_t>(b) * p.Sq + q) * p.Sk : nullptr;
      float m = -INFINITY, l = 0.f;  // running row maximum (shared by the row's two threads), this thread's half-row sum

      for (int j = 0; j < nblk; ++j, ++s) {
        const uint32_t ph = s & 1;
        const int k0 = j * kBK;
        float* kb = sKb + (s & 1) * kBK;
        if (threadIdx.x < kBK) {
          const int k = k0 + static_cast<int>(threadIdx.x);
          float v = -INFINITY;  // keys past the sequence end
          if (k < p.Sk) {
            const bool valid = (p.key_valid == nullptr) || __ldg(p.key_valid + static_cast<size_t>(b) * p.Sk + k);
            v = (valid ? 0.f : mask2) + (kPair ? pb2 : 0.f);
          }
          kb[threadIdx.x] = v;
        }
        if (threadIdx.x == 0) tma_store_wait_read();  // the previous item's output tile has left the P buffer
        named_bar_sync(1, kMathThreads);              // key bias visible; P buffer free
        mbar_wait(s_ready, ph);
        tc_fence_after();
        const uint32_t tSr = tS + lane_sel + wq * 64;
        const float* kbh = kb + wq * 64;
        const int key_h = k0 + wq * 64;
        // ---- pass A: row maximum over this thread's 64 keys
        float mloc = -INFINITY;
        if (warp_live) {
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            uint32_t vs[32];
            tmem_ld32(tSr + sub * 32, vs);
            tmem_ld_wait();
            bias32(vs, kbh + sub * 32, pair_row, key_h + sub * 32);
#pragma unroll
            for (int i = 0; i < 32; ++i) mloc = fmaxf(mloc, __uint_as_float(vs[i]));
          }
        }
        sMax[wq * kBQ + r] = mloc;
        named_bar_sync(2, kMathThreads);
        const float m_blk = fmaxf(sMax[r], sMax[kBQ + r]);
        const float m_new = fmaxf(m, m_blk);
        const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
        const float alpha = ex2a(m - m_use);  // m = -inf on the first block -> 0
        if (j > 0) {
          // the running output (this thread's 32 columns, in TMEM) takes the new maximum before this block is added;
          // P.V of the previous block must have landed
          mbar_wait(o_ready, ph ^ 1);
          tc_fence_after();
          if (warp_live && __any_sync(0xffffffffu, alpha != 1.0f)) {   // warp-uniform: tcgen05.ld / st are warp collectives
            uint32_t v[32];
            tmem_ld32(tO + lane_sel + wq * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st32(tO + lane_sel + wq * 32, v);
            tmem_st_wait();
          }
          tc_fence_before();
        }
        // ---- pass B: P = exp2(score - m), dropout, bf16 into the swizzled P panel (wq) of 64 keys
        float lsum = 0.f;
        uint8_t* prow = sP + wq * 16384 + r * 128;
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          if (!warp_live) break;
          uint32_t vs[32];
          tmem_ld32(tSr + sub * 32, vs);
          tmem_ld_wait();
          bias32(vs, kbh + sub * 32, pair_row, key_h + sub * 32);
          float e[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            e[i] = ex2a(__uint_as_float(vs[i]) - m_use);
            lsum += e[i];
          }
          if (p.drop.thr) {
            // dropout acts on the normalised probabilities; the row sum keeps the undropped values, the P that
            // multiplies V carries mask / (1 - p)
            const uint32_t e0 = static_cast<uint32_t>(((static_cast<size_t>(b) * p.heads + h) * p.Sq + q) * p.Sk + key_h + sub * 32);
            if ((p.Sk & 1) == 0) {
#pragma unroll
              for (int i = 0; i < 32; i += 2) {
                float m0, m1;
                drop_mul2(p.drop, e0 + i, m0, m1);
                e[i] *= m0; e[i + 1] *= m1;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) e[i] *= drop_mul(p.drop, e0 + i);
            }
          }
#pragma unroll
          for (int g = 0; g < 4; ++g)
            *reinterpret_cast<uint4*>(prow + (((sub * 4 + g) ^ (r & 7)) << 4)) = pack8f(e + 8 * g);
        }
        tc_fence_before();  // S fully read: the control thread may overwrite it once p_ready completes
        l = l * alpha + lsum;
        m = m_new;
        fence_proxy_async();  // make the generic-proxy smem writes visible to the tensor core (async proxy)
        mbar_arrive(p_ready);
      }
      // ---- the item's output: P.V of the last block has landed
      mbar_wait(o_ready, (s - 1) & 1);
      tc_fence_after();
      float o[32];
      if (warp_live) {
        uint32_t v[32];
        tmem_ld32(tO + lane_sel + wq * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(v[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0.f;
      }
      tc_fence_before();
      // ---- item epilogue: row sum over the two halves, normalise, stage the bf16 tile, one TMA store, lse ----
      sL[wq * kBQ + r] = l;
      named_bar_sync(2, kMathThreads);
      const float lt = sL[r] + sL[kBQ + r];
      const float inv = 1.0f / lt;
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i] *= inv;
      {
        uint8_t* orow = sP + r * 128;  // (P.V of the last step is complete: o_ready was observed by every thread)
#pragma unroll
        for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(orow + (((wq * 4 + g) ^ (r & 7)) << 4)) = pack8f(o + 8 * g);
      }
      if (wq == 0 && q < p.Sq && p.lse) p.lse[(static_cast<size_t>(b) * p.heads + h) * p.Sq + q] = (m + log2f(lt)) * kLn2;
      fence_proxy_async();
      named_bar_sync(1, kMathThreads);  // tile complete
      if (threadIdx.x == 0) {
        tma_store_3d(&tmO, sP, h * kD, q0, b);  // rows past Sq are clipped by the TMA unit
        tma_store_commit();
      }
    }
    if (threadIdx.x == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMathWarps) {
    tc_fence_after();
    tmem_dealloc(tS, kTmemCols);
  }
}

}  // namespace

bool attention_tc_supported(const AttnArgs& a);  // attention_tc.cu

int attention_tc2_fwd(const AttnArgs& a, cudaStream_t stream) {
  ETP_REQUIRE(a.B > 0 && a.Sq > 0 && a.Sk > 0 && a.heads > 0, "attention_tc2: empty problem");
  ETP_REQUIRE(attention_tc_supported(a), "attention_tc2: unsupported layout");
  CUtensorMap tq, tk, tv, to;
  const uint64_t W = static_cast<uint64_t>(a.heads) * kD;
  int rc = get_tmap_3d(a.q, W, a.Sq, a.B, a.ldq, static_cast<uint64_t>(a.Sq) * a.ldq, kD, kBQ, &tq);
  if (rc) return rc;
  const int kvB = a.kv_rows ? a.kv_B : a.B;
  ETP_REQUIRE(kvB > 0, "attention_tc2: kv_B must be given with kv_rows");
  rc = get_tmap_3d(a.k, W, a.Sk, kvB, a.ldk, static_cast<uint64_t>(a.Sk) * a.ldk, kD, kBK, &tk);
  if (rc) return rc;
  rc = get_tmap_3d(a.v, W, a.Sk, kvB, a.ldv, static_cast<uint64_t>(a.Sk) * a.ldv, kD, kBK, &tv);
  if (rc) return rc;
  rc = get_tmap_3d(a.out, W, a.Sq, a.B, a.ldo, static_cast<uint64_t>(a.Sq) * a.ldo, kD, kBQ, &to);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    ETP_CHECK_CUDA(cudaFuncSetAttribute(attention_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    ETP_CHECK_CUDA(cudaFuncSetAttribute(attention_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  Attn2Dev d;
  d.B = a.B; d.heads = a.heads; d.Sq = a.Sq; d.Sk = a.Sk; d.scale = a.scale; d.key_valid = a.key_valid;
  d.mask_value = a.mask_value; d.pair = a.pair; d.pair_w = a.pair_w; d.pair_b = a.pair_b;
  d.pair_w_dev = a.pair_w_dev; d.pair_b_dev = a.pair_b_dev; d.lse = a.lse;
  d.drop = Drop{a.drop_key, a.drop_thr, a.drop_scale};
  d.kv_rows = a.kv_rows;
  ETP_REQUIRE(!a.drop_thr || static_cast<int64_t>(a.B) * a.heads * a.Sq * a.Sk < (int64_t(1) << 32), "attention: dropout index range");
  const int items = a.B * a.heads * ((a.Sq + kBQ - 1) / kBQ);
  const int slots = 2 * num_sms();  // two resident CTAs per SM
  const int grid = items < slots ? items : slots;
  if (a.pair)
    ETP_CHECK_CUDA(launch_pdl(attention_tc2_kernel<true>, dim3(grid), dim3(kThreads), kSmemBytes, stream, tq, tk, tv, to, d));
  else
    ETP_CHECK_CUDA(launch_pdl(attention_tc2_kernel<false>, dim3(grid), dim3(kThreads), kSmemBytes, stream, tq, tk, tv, to, d));
  ETP_LAUNCHED();
  return ETP_OK;
}

// 1 (default): two-CTAs-per-SM kernel of this file; 0: the one-CTA kernel of attention_tc.cu (A/B measurements)
bool attention_use_v2() {
  static const int v = [] { const char* e = getenv("ETP_ATTN_V2"); return e ? atoi(e) : 1; }();
  return v != 0;
}

}  // namespace etp
