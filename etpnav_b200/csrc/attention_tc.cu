// Multi-head attention (head dim 64) on tcgen05 tensor cores: text<->graph cross-attention and the
// graph-aware node self-attention of the planner (BertOutAttention vilmodel_cmt.py:325-352,
// BertSelfAttention :103-141 with the sprel bias of :391-393), and the language encoder's self-attention.
//
// One CTA = one (batch, head, 128-query tile).  Q, and K/V in blocks of 128 keys, are brought in by TMA
// (3-D tensor maps over [B, S, heads*64], 128-byte swizzle; rows past the sequence end are zero-filled by
// the TMA unit).  Per key block:   S = Q.K^T  (tcgen05.mma M=128 N=128 K=64, fp32 in TMEM)
//   -> four softmax warps, one query row per thread.  Pass 1: tcgen05.ld the scores, add the per-key mask
//      (shared-memory broadcast) and the pair bias (staged coalesced through shared memory, 32 keys at a
//      time), all in the log2 domain, track the row max and write the biased scores back to TMEM
//      (tcgen05.st).  Pass 2: tcgen05.ld, P = exp2(s - m) -> bf16 -> shared memory in the K-major
//      128B-swizzled layout the tensor core reads; online row sum;
//   -> O_blk = P.V (tcgen05.mma M=128 N=64 K=128; V is consumed MN-major exactly as it lies in memory),
//      read back and folded into the per-thread fp32 running output with the usual rescale.
// ~98 KB shared memory and 256 TMEM columns per CTA, so two CTAs share an SM and hide each other's
// load / softmax phases.
#include "common.cuh"
#include "host.h"
#include "ops.h"

namespace etp {

namespace {

constexpr int kBQ = 128;   // query rows per CTA
constexpr int kBK = 128;   // keys per block
constexpr int kD = 64;
constexpr int kQBytes = kBQ * kD * 2;       // 16 KB
constexpr int kKBytes = kBK * kD * 2;       // 16 KB
constexpr int kPBytes = kBQ * kBK * 2;      // 32 KB (two 64-key swizzle panels of 16 KB)
constexpr int kPairStride = 33;
constexpr int kPairBytes = kBQ * kPairStride * 4;
constexpr int kSmemBytes = kQBytes + 2 * kKBytes + kPBytes + kPairBytes + kBK * 4 + 1024 + 256;
constexpr int kThreads = 160;               // 4 softmax warps + 1 control warp
constexpr uint32_t kTmemCols = 256;         // S: [0,128)  O_blk: [128,192)
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
  bf16* out;
  int ldo;
  float* lse;
};

__global__ void __launch_bounds__(kThreads, 2)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kQBytes;
  uint8_t* sV = sK + kKBytes;
  uint8_t* sP = sV + kKBytes;
  float* sPair = reinterpret_cast<float*>(sP + kPBytes);
  float* sKb = sPair + kBQ * kPairStride;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKb + kBK);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* v_full = bars + 2;
  uint64_t* s_ready = bars + 3;
  uint64_t* p_ready = bars + 4;
  uint64_t* o_ready = bars + 5;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kBQ;
  const int nblk = (p.Sk + kBK - 1) / kBK;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(v_full, 1);
    mbar_init(s_ready, 1);
    mbar_init(p_ready, 128);
    mbar_init(o_ready, 1);
    fence_barrier_init();
    fence_proxy_async();
  }
  griddep_launch();  // PDL: the next kernel may start its own prologue
  griddep_wait();    // previous kernel complete; nothing above touched global memory or TMEM
  if (warp == 4) {
    tmem_alloc(tmem_ptr, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tmem_S = tmem_base;
  const uint32_t tmem_O = tmem_base + 128;

  if (warp == 4) {
    // ======================= control warp: TMA + MMA issue (one lane) =======================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(kBQ, kBK, 0, 0);  // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_o = make_idesc_bf16(kBQ, kD, 0, 1);   // P (K-major) x V (MN-major)
      mbar_arrive_expect_tx(q_full, kQBytes);
      tma_load_3d(sQ, &tmQ, q_full, h * kD, q0, b);
      for (int j = 0; j < nblk; ++j) {
        const uint32_t ph = j & 1;
        if (j > 0) mbar_wait(o_ready, ph ^ 1);  // P.V of block j-1 done: K, V, P buffers are free again
        mbar_arrive_expect_tx(k_full, kKBytes);
        tma_load_3d(sK, &tmK, k_full, h * kD, j * kBK, b);
        mbar_arrive_expect_tx(v_full, kKBytes);
        tma_load_3d(sV, &tmV, v_full, h * kD, j * kBK, b);
        if (j == 0) mbar_wait(q_full, 0);
        mbar_wait(k_full, ph);
        tc_fence_after();
        const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k)
          umma_bf16(tmem_S, make_smem_desc(aq + k * 32, 16, 1024), make_smem_desc(ak + k * 32, 16, 1024), idesc_s,
                    k > 0 ? 1u : 0u);
        umma_commit(s_ready);
        mbar_wait(p_ready, ph);
        mbar_wait(v_full, ph);
        tc_fence_after();
        const uint32_t ap = smem_u32(sP), av = smem_u32(sV);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k)
          umma_bf16(tmem_O, make_smem_desc(ap + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                    make_smem_desc(av + k * 2048, 8192, 1024), idesc_o, k > 0 ? 1u : 0u);
        umma_commit(o_ready);
      }
    }
  } else {
    // ======================= softmax warps: one query row per thread =======================
    const int r = threadIdx.x;          // row in the tile == TMEM lane
    const int q = q0 + r;
    const bool qv = q < p.Sq;
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    const float pw = (p.pair_w_dev ? __ldg(p.pair_w_dev) : p.pair_w) * kLog2e;
    const float pb = (p.pair_b_dev ? __ldg(p.pair_b_dev) : p.pair_b) * kLog2e;
    const float sl2 = p.scale * kLog2e;
    const float mask2 = p.mask_value * kLog2e;
    const uint8_t* kvalid = p.key_valid ? p.key_valid + static_cast<size_t>(b) * p.Sk : nullptr;
    const float* pair_b0 = p.pair ? p.pair + static_cast<size_t>(b) * p.Sq * p.Sk : nullptr;
    float m = -INFINITY, l = 0.f;
    float o[kD];
#pragma unroll
    for (int i = 0; i < kD; ++i) o[i] = 0.f;

    for (int j = 0; j < nblk; ++j) {
      const uint32_t ph = j & 1;
      const int k0 = j * kBK;
      // per-key additive mask of this block (log2 domain); keys past the sequence end get -inf
      {
        const int k = k0 + r;
        float kb = -INFINITY;
        if (k < p.Sk) kb = (kvalid && !kvalid[k]) ? mask2 : 0.f;
        sKb[r] = kb;
      }
      // first 32-key chunk of the pair bias: issue the (coalesced) loads now, they land while the MMA runs
      float pr[32];
      auto load_pair = [&](int c) {
        const int key = k0 + c + lane;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int row = warp + 4 * i;
          pr[i] = (q0 + row < p.Sq && key < p.Sk) ? __ldg(pair_b0 + static_cast<size_t>(q0 + row) * p.Sk + key) : 0.f;
        }
      };
      if (pair_b0) load_pair(0);
      named_bar_sync(1, 128);
      mbar_wait(s_ready, ph);
      tc_fence_after();
      // pass 1: biased scores (log2 domain) -> TMEM, block max
      float mj = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < kBK; c += 32) {
        if (pair_b0) {
          // coalesced staging of pair[b, q0 + row, k0 + c + lane] for the 128 rows of the tile
#pragma unroll
          for (int i = 0; i < 32; ++i) sPair[(warp + 4 * i) * kPairStride + lane] = pr[i];
          named_bar_sync(1, 128);
          if (c + 32 < kBK) load_pair(c + 32);  // software prefetch of the next chunk
        }
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_sel + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float s = fmaf(__uint_as_float(v[i]), sl2, sKb[c + i]);
          if (pair_b0) s += fmaf(pw, sPair[r * kPairStride + i], pb);
          mj = fmaxf(mj, s);
          v[i] = __float_as_uint(s);
        }
        tmem_st32(tmem_S + lane_sel + c, v);
        if (pair_b0) named_bar_sync(1, 128);  // staging buffer is reused by the next chunk
      }
      tmem_st_wait();
      const float m_new = fmaxf(m, mj);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = exp2f(m - m_use);  // m = -inf on the first block -> 0
      // pass 2: P = exp2(s - m) -> bf16 -> swizzled shared memory; row sum
      float lsum = 0.f;
#pragma unroll 1
      for (int c = 0; c < kBK; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_S + lane_sel + c, v);
        tmem_ld_wait();
        float e[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          e[i] = exp2f(__uint_as_float(v[i]) - m_use);
          lsum += e[i];
        }
        // 32 keys = 4 chunks of 16 B inside the 64-key panel (c >> 6)
        uint8_t* prow_s = sP + (c >> 6) * 16384 + r * 128;
        const int ch0 = (c & 63) >> 3;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 u = make_uint4(pack_bf16x2(e[8 * g], e[8 * g + 1]), pack_bf16x2(e[8 * g + 2], e[8 * g + 3]),
                                     pack_bf16x2(e[8 * g + 4], e[8 * g + 5]), pack_bf16x2(e[8 * g + 6], e[8 * g + 7]));
          *reinterpret_cast<uint4*>(prow_s + (((ch0 + g) ^ (r & 7)) << 4)) = u;
        }
      }
      l = l * alpha + lsum;
      m = m_new;
      fence_proxy_async();  // make the generic-proxy smem writes visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(p_ready);
      // fold O_blk into the running output
      mbar_wait(o_ready, ph);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < kD; c += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_O + lane_sel + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[c + i] = o[c + i] * alpha + __uint_as_float(v[i]);
      }
      tc_fence_before();
    }
    if (qv) {
      const float inv = 1.0f / l;
      bf16* og = p.out + (static_cast<size_t>(b) * p.Sq + q) * p.ldo + h * kD;
#pragma unroll
      for (int i = 0; i < kD; i += 8) {
        const uint4 u = make_uint4(pack_bf16x2(o[i] * inv, o[i + 1] * inv), pack_bf16x2(o[i + 2] * inv, o[i + 3] * inv),
                                   pack_bf16x2(o[i + 4] * inv, o[i + 5] * inv), pack_bf16x2(o[i + 6] * inv, o[i + 7] * inv));
        *reinterpret_cast<uint4*>(og + i) = u;
      }
      if (p.lse) p.lse[(static_cast<size_t>(b) * p.heads + h) * p.Sq + q] = (m + log2f(l)) * kLn2;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace

bool attention_tc_supported(const AttnArgs& a) {
  if (a.ldq % 8 || a.ldk % 8 || a.ldv % 8 || a.ldo % 8) return false;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return al16(a.q) && al16(a.k) && al16(a.v) && al16(a.out);
}

int attention_tc_fwd(const AttnArgs& a, cudaStream_t stream) {
  ETP_REQUIRE(a.B > 0 && a.Sq > 0 && a.Sk > 0 && a.heads > 0, "attention_tc: empty problem");
  ETP_REQUIRE(attention_tc_supported(a), "attention_tc: unsupported layout");
  CUtensorMap tq, tk, tv;
  const uint64_t W = static_cast<uint64_t>(a.heads) * kD;
  int rc = get_tmap_3d(a.q, W, a.Sq, a.B, a.ldq, static_cast<uint64_t>(a.Sq) * a.ldq, kD, kBQ, &tq);
  if (rc) return rc;
  rc = get_tmap_3d(a.k, W, a.Sk, a.B, a.ldk, static_cast<uint64_t>(a.Sk) * a.ldk, kD, kBK, &tk);
  if (rc) return rc;
  rc = get_tmap_3d(a.v, W, a.Sk, a.B, a.ldv, static_cast<uint64_t>(a.Sk) * a.ldv, kD, kBK, &tv);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    ETP_CHECK_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  AttnDev d;
  d.B = a.B; d.heads = a.heads; d.Sq = a.Sq; d.Sk = a.Sk; d.scale = a.scale; d.key_valid = a.key_valid;
  d.mask_value = a.mask_value; d.pair = a.pair; d.pair_w = a.pair_w; d.pair_b = a.pair_b;
  d.pair_w_dev = a.pair_w_dev; d.pair_b_dev = a.pair_b_dev; d.out = a.out; d.ldo = a.ldo; d.lse = a.lse;
  dim3 grid((a.Sq + kBQ - 1) / kBQ, a.heads, a.B);
  ETP_CHECK_CUDA(launch_pdl(attention_tc_kernel, dim3(grid), dim3(kThreads), kSmemBytes, stream, tq, tk, tv, d));
  ETP_LAUNCHED();
  return ETP_OK;
}

}  // namespace etp
