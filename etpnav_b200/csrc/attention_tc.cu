// Multi-head attention (head dim 64) on tcgen05 tensor cores: text<->graph cross-attention and the
// graph-aware node self-attention of the planner (BertOutAttention vilmodel_cmt.py:325-352,
// BertSelfAttention :103-141 with the sprel bias of :391-393), and the language encoder's self-attention.
//
// PERSISTENT: one CTA per SM walks over (batch, head, 128-query tile) items; keys in blocks of 128 ("steps").
// Q, K, V tiles arrive by TMA (3-D tensor maps over [B, S, heads*64], 128-byte swizzle, rows past the sequence
// end zero-filled), double-buffered so the loads of step s+1 fly while step s is computed.
//   control thread:  S = Q.K^T  (tcgen05.mma M=128 N=128 K=64, fp32 in TMEM, two S buffers: S of step s+1 is
//                    issued right behind P.V of step s)
//   16 softmax warps, thread = (query row, 32-key slice): scores -> registers (one TMEM read, no write-back),
//                    + key mask + pair bias in the log2 domain, slice max -> row max through shared memory,
//                    P = exp2(s - m) -> bf16 -> shared memory in the K-major 128B-swizzled layout the MMA reads;
//   control thread:  O_blk = P.V (M=128 N=64 K=128; V consumed MN-major exactly as it lies in memory);
//   softmax warps:   each thread folds 16 of the 64 output columns into its running fp32 output (usual rescale).
// The finished tile is normalised, written as bf16 into shared memory (TMA swizzled layout) and leaves with ONE
// TMA store (row-per-thread global stores would touch 32 cache lines per warp instruction); lse goes out
// coalesced.  TMEM allocation, barrier set-up and launch are paid once per CTA.
#include "common.cuh"
#include "host.h"
#include "ops.h"

namespace etp {

namespace {

constexpr int kBQ = 128;   // query rows per item
constexpr int kBK = 128;   // keys per step
constexpr int kD = 64;
constexpr int kTile = kBQ * kD * 2;      // 16 KB: Q, K, V tiles
constexpr int kPBytes = kBQ * kBK * 2;   // 32 KB (two 64-key swizzle panels of 16 KB)
constexpr int kWQ = 4;                   // warps per TMEM lane quadrant: each owns 32 of the step's 128 keys
constexpr int kWarps = 4 * kWQ;
constexpr int kMathThreads = kWarps * 32;
constexpr int kThreads = kMathThreads + 32;  // + control warp
constexpr int kSmemBytes = 6 * kTile + kPBytes + 2 * kBK * 4 + 2 * kWQ * kBQ * 4 + 256 + 1024 + 256;
constexpr uint32_t kTmemCols = 512;      // S0 [0,128)  S1 [128,256)  O_blk [256,320)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct AttnDev {
  int B, heads, Sq, Sk;
  float scale;
  const uint8_t* key_valid;
  float mask_value;
  const float* pair;
  float pair_w, pair_b;
  const float* pair_w_dev;
  const float* pair_b_dev;
  float* lse;
  Drop drop;  // attention-probability dropout (thr 0 = off)
  const int32_t* kv_rows;  // optional batch-row map of the K / V tensor (see AttnArgs)
};

ETP_DEVICE float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
ETP_DEVICE uint4 pack8(const float* f) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

template <bool kPair>
__global__ void __launch_bounds__(kThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const AttnDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;               // [2][16 KB]
  uint8_t* sK = sQ + 2 * kTile;     // [2][16 KB]
  uint8_t* sV = sK + 2 * kTile;     // [2][16 KB]
  uint8_t* sP = sV + 2 * kTile;     // 32 KB; its first 16 KB double as the output tile of the TMA store
  float* sKb = reinterpret_cast<float*>(sP + kPBytes);  // [2][128] per-key additive bias (log2 domain)
  float* sMax = sKb + 2 * kBK;                           // [kWQ][128] slice maxima of the current step
  float* sL = sMax + kWQ * kBQ;                          // [kWQ][128] slice row sums at item end
  uint64_t* bars = reinterpret_cast<uint64_t*>(sL + kWQ * kBQ);
  uint64_t* q_full = bars + 0;    // [2]
  uint64_t* kv_full = bars + 2;   // [2]
  uint64_t* s_ready = bars + 4;   // [2]
  uint64_t* p_ready = bars + 6;
  uint64_t* o_ready = bars + 7;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = (p.Sk + kBK - 1) / kBK;
  const int qtiles = (p.Sq + kBQ - 1) / kBQ;
  const int n_items = p.B * p.heads * qtiles;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmO);
    mbar_init(&q_full[0], 1); mbar_init(&q_full[1], 1);
    mbar_init(&kv_full[0], 1); mbar_init(&kv_full[1], 1);
    mbar_init(&s_ready[0], 1); mbar_init(&s_ready[1], 1);
    mbar_init(p_ready, kMathThreads);
    mbar_init(o_ready, 1);
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
  const uint32_t tO = tm + 256;

  // item -> (batch, head, query tile); query tiles of one (b, h) are adjacent items, i.e. run on neighbouring SMs
  auto item_bhq = [&](int item, int& b, int& h, int& q0) {
    const int qt = item % qtiles;
    const int bh = item / qtiles;
    b = bh / p.heads; h = bh % p.heads; q0 = qt * kBQ;
  };

  if (warp == kWarps) {
    // ======================= control warp: TMA + MMA issue (one lane) =======================
    if (lane == 0 && static_cast<int>(blockIdx.x) < n_items) {
      constexpr uint32_t idesc_s = make_idesc_bf16(kBQ, kBK, 0, 0);  // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = make_idesc_bf16(kBQ, kD, 0, 1);   // P (K-major) x V (MN-major)
      const uint64_t dp0 = make_smem_desc(smem_u32(sP), 16, 1024);
      auto load_q = [&](int item, int st) {
        int b, h, q0;
        item_bhq(item, b, h, q0);
        mbar_arrive_expect_tx(&q_full[st], kTile);
        tma_load_3d(sQ + st * kTile, &tmQ, &q_full[st], h * kD, q0, b);
      };
      auto load_kv = [&](int item, int j, int st) {
        int b, h, q0;
        item_bhq(item, b, h, q0);
        const int kvb = p.kv_rows ? __ldg(p.kv_rows + b) : b;
        mbar_arrive_expect_tx(&kv_full[st], 2 * kTile);
        tma_load_3d(sK + st * kTile, &tmK, &kv_full[st], h * kD, j * kBK, kvb);
        tma_load_3d(sV + st * kTile, &tmV, &kv_full[st], h * kD, j * kBK, kvb);
      };
      // S of step (ii, s): waits for its tiles, then 4 MMAs into S buffer s & 1
      auto issue_s = [&](int ii, int s, bool first_block) {
        if (first_block) mbar_wait(&q_full[ii & 1], (ii >> 1) & 1);
        mbar_wait(&kv_full[s & 1], (s >> 1) & 1);
        tc_fence_after();
        const uint64_t dq0 = make_smem_desc(smem_u32(sQ + (ii & 1) * kTile), 16, 1024);
        const uint64_t dk0 = make_smem_desc(smem_u32(sK + (s & 1) * kTile), 16, 1024);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          umma_bf16_lh(tm + (s & 1) * 128, desc_lo(dq0) + 2 * k, desc_hi(dq0), desc_lo(dk0) + 2 * k, desc_hi(dk0), idesc_s,
                       k > 0 ? 1u : 0u);
        umma_commit(&s_ready[s & 1]);
      };
      load_q(blockIdx.x, 0);
      load_kv(blockIdx.x, 0, 0);
      issue_s(0, 0, true);
      int s = 0, ii = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++ii) {
        for (int j = 0; j < nblk; ++j, ++s) {
          // tiles of step s+1: its K/V stage was last read by P.V of step s-1
          const bool more_blocks = j + 1 < nblk;
          const int next_item = item + gridDim.x;
          const bool has_next = more_blocks || next_item < n_items;
          if (has_next) {
            if (s >= 1) mbar_wait(o_ready, (s - 1) & 1);
            if (more_blocks) {
              load_kv(item, j + 1, (s + 1) & 1);
            } else {
              load_q(next_item, (ii + 1) & 1);
              load_kv(next_item, 0, (s + 1) & 1);
            }
          }
          // O_blk = P.V of this step
          mbar_wait(p_ready, s & 1);
          tc_fence_after();
          const uint64_t dv0 = make_smem_desc(smem_u32(sV + (s & 1) * kTile), 8192, 1024);
          // only the key groups that hold real keys (P is 0 beyond them): every MMA costs ~75 ns of issue latency
          const int kgroups = (min(kBK, p.Sk - j * kBK) + 15) >> 4;
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            if (k < kgroups)
            umma_bf16_lh(tO, desc_lo(dp0) + (k >> 2) * 1024 + (k & 3) * 2, desc_hi(dp0), desc_lo(dv0) + 128 * k, desc_hi(dv0),
                         idesc_o, k > 0 ? 1u : 0u);
          umma_commit(o_ready);
          // S of the next step runs behind it (its S buffer was consumed before p_ready of step s-1)
          if (has_next) issue_s(more_blocks ? ii : ii + 1, s + 1, !more_blocks);
        }
      }
    }
  } else {
    // ======================= softmax warps: thread = (query row, 32-key slice) =======================
    const int quad = warp & 3;       // TMEM lane quadrant
    const int wq = warp >> 2;        // which 32 keys of the step / which 16 output columns
    const int r = quad * 32 + lane;  // row in the tile == TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(quad * 32) << 16;
    const float pw2 = (p.pair_w_dev ? __ldg(p.pair_w_dev) : p.pair_w) * kLog2e;
    const float pb2 = (p.pair_b_dev ? __ldg(p.pair_b_dev) : p.pair_b) * kLog2e;
    const float sl2 = p.scale * kLog2e;
    const float mask2 = p.mask_value * kLog2e;
    const bool pair_vec = (p.Sk & 3) == 0;  // rows of the pair bias are 16-byte aligned

    // per-step inputs fetched one step ahead: validity of key (k0 + tid) and this thread's first 16 pair biases
    uint8_t kv_next = 1;
    float pv[16];
    auto load_pair_at = [&](int b, int q, int key0) {
      if constexpr (kPair) {
        if (q < p.Sq) {
          const float* row_ptr = p.pair + (static_cast<size_t>(b) * p.Sq + q) * p.Sk;
          if (pair_vec) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
              if (key0 + i < p.Sk) f = __ldg(reinterpret_cast<const float4*>(row_ptr + key0 + i));
              pv[i] = f.x; pv[i + 1] = f.y; pv[i + 2] = f.z; pv[i + 3] = f.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) pv[i] = (key0 + i < p.Sk) ? __ldg(row_ptr + key0 + i) : 0.f;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) pv[i] = 0.f;
        }
      }
    };
    auto request_step = [&](int item, int j) {
      if (item >= n_items) return;
      int b, h, q0;
      item_bhq(item, b, h, q0);
      if (threadIdx.x < kBK) {
        const int k = j * kBK + static_cast<int>(threadIdx.x);
        kv_next = (p.key_valid && k < p.Sk) ? __ldg(p.key_valid + static_cast<size_t>(b) * p.Sk + k) : 1;
      }
      load_pair_at(b, q0 + r, j * kBK + wq * 32);
    };
    request_step(blockIdx.x, 0);

    int s = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int b, h, q0;
      item_bhq(item, b, h, q0);
      const int q = q0 + r;
      float m = -INFINITY, l = 0.f;  // running row max (shared by the row's 4 threads) and this thread's slice sum
      float o[16];                   // this thread's 16 output columns
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] = 0.f;

      for (int j = 0; j < nblk; ++j, ++s) {
        const uint32_t ph = s & 1;
        const int k0 = j * kBK;
        float* kb = sKb + (s & 1) * kBK;
        if (threadIdx.x < kBK) {
          const int k = k0 + static_cast<int>(threadIdx.x);
          float v = -INFINITY;  // keys past the sequence end
          if (k < p.Sk) v = (kv_next ? 0.f : mask2) + (kPair ? pb2 : 0.f);
          kb[threadIdx.x] = v;
        }
        if (threadIdx.x == 0) tma_store_wait_read();  // the previous item's output tile has left the P buffer
        named_bar_sync(1, kMathThreads);              // key bias visible; P buffer free
        mbar_wait(&s_ready[s & 1], (s >> 1) & 1);
        tc_fence_after();
        // scores of this thread's 32 keys, biased, log2 domain
        float sc[32];
        float mloc = -INFINITY;
        const uint32_t tS = tm + (s & 1) * 128 + lane_sel + wq * 32;
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          uint32_t vs[16];
          tmem_ld16(tS + sub * 16, vs);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const float4 kb4 = *reinterpret_cast<const float4*>(kb + wq * 32 + sub * 16 + i);
            const float kbv[4] = {kb4.x, kb4.y, kb4.z, kb4.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              float x = fmaf(__uint_as_float(vs[i + t]), sl2, kbv[t]);
              if constexpr (kPair) x = fmaf(pw2, pv[i + t], x);
              sc[sub * 16 + i + t] = x;
              mloc = fmaxf(mloc, x);
            }
          }
          if (sub == 0) load_pair_at(b, q, k0 + wq * 32 + 16);  // second half of this step's bias
        }
        tc_fence_before();
        sMax[wq * kBQ + r] = mloc;
        named_bar_sync(2, kMathThreads);
        const float m_blk = fmaxf(fmaxf(sMax[r], sMax[kBQ + r]), fmaxf(sMax[2 * kBQ + r], sMax[3 * kBQ + r]));
        const float m_new = fmaxf(m, m_blk);
        const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
        const float alpha = ex2_approx(m - m_use);  // m = -inf on the first block -> 0
        // next step's mask / bias: in flight while P is computed
        if (j + 1 < nblk) request_step(item, j + 1);
        else request_step(item + gridDim.x, 0);
        float lsum = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          sc[i] = ex2_approx(sc[i] - m_use);
          lsum += sc[i];
        }
        if (p.drop.thr) {
          // dropout acts on the normalised probabilities; the row sum keeps the undropped values, the P that
          // multiplies V carries mask / (1 - p)
          const uint32_t e0 = static_cast<uint32_t>(((static_cast<size_t>(b) * p.heads + h) * p.Sq + q) * p.Sk + k0 + wq * 32);
          if ((p.Sk & 1) == 0) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float m0, m1;
              drop_mul2(p.drop, e0 + i, m0, m1);
              sc[i] *= m0; sc[i + 1] *= m1;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) sc[i] *= drop_mul(p.drop, e0 + i);
          }
        }
        {
          // 32 keys = 4 chunks of 16 B inside the 64-key panel (wq >> 1)
          uint8_t* prow = sP + (wq >> 1) * 16384 + r * 128;
          const int ch0 = (wq & 1) * 4;
#pragma unroll
          for (int g = 0; g < 4; ++g) *reinterpret_cast<uint4*>(prow + (((ch0 + g) ^ (r & 7)) << 4)) = pack8(sc + 8 * g);
        }
        l = l * alpha + lsum;
        m = m_new;
        fence_proxy_async();  // make the generic-proxy smem writes visible to the tensor core (async proxy)
        mbar_arrive(p_ready);
        // fold this thread's 16 columns of O_blk into the running output
        mbar_wait(o_ready, ph);
        tc_fence_after();
        {
          uint32_t v[16];
          tmem_ld16(tO + lane_sel + wq * 16, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = fmaf(o[i], alpha, __uint_as_float(v[i]));
        }
        tc_fence_before();
      }
      // ---- item epilogue: row sum over the 4 slices, normalise, stage the bf16 tile, one TMA store, lse ----
      sL[wq * kBQ + r] = l;
      named_bar_sync(2, kMathThreads);
      const float lt = sL[r] + sL[kBQ + r] + sL[2 * kBQ + r] + sL[3 * kBQ + r];
      const float inv = 1.0f / lt;
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] *= inv;
      {
        uint8_t* orow = sP + r * 128;  // (P.V of the last step is complete: o_ready was observed by every thread)
#pragma unroll
        for (int g = 0; g < 2; ++g) *reinterpret_cast<uint4*>(orow + (((wq * 2 + g) ^ (r & 7)) << 4)) = pack8(o + 8 * g);
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
  if (warp == kWarps) {
    tc_fence_after();
    tmem_dealloc(tm, kTmemCols);
  }
}

}  // namespace

bool attention_tc_supported(const AttnArgs& a) {
  if (a.ldq % 8 || a.ldk % 8 || a.ldv % 8 || a.ldo % 8) return false;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return al16(a.q) && al16(a.k) && al16(a.v) && al16(a.out) && (a.pair == nullptr || al16(a.pair));
}

int attention_tc2_fwd(const AttnArgs& a, cudaStream_t stream);  // attention_tc2.cu: two CTAs per SM
bool attention_use_v2();

int attention_tc_fwd_v1(const AttnArgs& a, cudaStream_t stream);
int attention_tc_fwd(const AttnArgs& a, cudaStream_t stream) {
  return attention_use_v2() ? attention_tc2_fwd(a, stream) : attention_tc_fwd_v1(a, stream);
}

int attention_tc_fwd_v1(const AttnArgs& a, cudaStream_t stream) {
  ETP_REQUIRE(a.B > 0 && a.Sq > 0 && a.Sk > 0 && a.heads > 0, "attention_tc: empty problem");
  ETP_REQUIRE(attention_tc_supported(a), "attention_tc: unsupported layout");
  CUtensorMap tq, tk, tv, to;
  const uint64_t W = static_cast<uint64_t>(a.heads) * kD;
  int rc = get_tmap_3d(a.q, W, a.Sq, a.B, a.ldq, static_cast<uint64_t>(a.Sq) * a.ldq, kD, kBQ, &tq);
  if (rc) return rc;
  const int kvB = a.kv_rows ? a.kv_B : a.B;
  ETP_REQUIRE(kvB > 0, "attention_tc: kv_B must be given with kv_rows");
  rc = get_tmap_3d(a.k, W, a.Sk, kvB, a.ldk, static_cast<uint64_t>(a.Sk) * a.ldk, kD, kBK, &tk);
  if (rc) return rc;
  rc = get_tmap_3d(a.v, W, a.Sk, kvB, a.ldv, static_cast<uint64_t>(a.Sk) * a.ldv, kD, kBK, &tv);
  if (rc) return rc;
  rc = get_tmap_3d(a.out, W, a.Sq, a.B, a.ldo, static_cast<uint64_t>(a.Sq) * a.ldo, kD, kBQ, &to);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    ETP_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    ETP_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  AttnDev d;
  d.B = a.B; d.heads = a.heads; d.Sq = a.Sq; d.Sk = a.Sk; d.scale = a.scale; d.key_valid = a.key_valid;
  d.mask_value = a.mask_value; d.pair = a.pair; d.pair_w = a.pair_w; d.pair_b = a.pair_b;
  d.pair_w_dev = a.pair_w_dev; d.pair_b_dev = a.pair_b_dev; d.lse = a.lse;
  d.drop = Drop{a.drop_key, a.drop_thr, a.drop_scale};
  d.kv_rows = a.kv_rows;
  ETP_REQUIRE(!a.drop_thr || static_cast<int64_t>(a.B) * a.heads * a.Sq * a.Sk < (int64_t(1) << 32), "attention: dropout index range");
  const int items = a.B * a.heads * ((a.Sq + kBQ - 1) / kBQ);
  const int grid = items < num_sms() ? items : num_sms();
  if (a.pair)
    ETP_CHECK_CUDA(launch_pdl(attention_tc_kernel<true>, dim3(grid), dim3(kThreads), kSmemBytes, stream, tq, tk, tv, to, d));
  else
    ETP_CHECK_CUDA(launch_pdl(attention_tc_kernel<false>, dim3(grid), dim3(kThreads), kSmemBytes, stream, tq, tk, tv, to, d));
  ETP_LAUNCHED();
  return ETP_OK;
}

}  // namespace etp
