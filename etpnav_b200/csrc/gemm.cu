// Persistent warp-specialised bf16 GEMM on tcgen05 tensor cores (sm_100a).
//
//   D[M,N] = epilogue( alpha * sum_k A[m,k] * B[n,k] )
//
//   * operands are staged tile-by-tile by TMA (cp.async.bulk.tensor, 128-byte swizzle) through a ring
//     of mbarrier-guarded shared-memory stages; one elected thread issues tcgen05.mma (M=128, N=BN,
//     K=16) with fp32 accumulators in TMEM; eight epilogue warps read the accumulator back with
//     tcgen05.ld and apply bias / GELU / ReLU / activation-derivative / residual and write fp32 and/or
//     bf16 outputs with 128-bit stores.  Two TMEM accumulator stages let the epilogue of tile i overlap
//     the main loop of tile i+1.
//   * each operand is either K-major (row-major [rows, K]) or MN-major ([K, rows], rows contiguous);
//     the latter serves dgrad (B = W as stored) and wgrad (A = dY, B = X as stored) without transposes.
//   * split-K with fp32 atomic accumulation serves wgrad (few output tiles, long reduction).
//
// Replaces the cuBLAS calls behind torch.nn.Linear / torch.matmul on the reference path
// (vilmodel_cmt.py:108-110,326-328,151,178,190,654; common/transformer.py:174-181).
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "gemm_dev.h"
#include "host.h"
#include "ops.h"

namespace etp {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + kEpiWarps * 32;

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator stages
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};


template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmDev p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full = empty_bar + Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpiWarps);
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  griddep_launch();  // PDL: the next kernel may start its own prologue
  griddep_wait();    // previous kernel complete; nothing above touched global memory or TMEM
  if (warp == 1) {
    tmem_alloc(tmem_ptr, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int num_tiles = p.tiles_m * p.tiles_n * p.k_splits;
  const int total_kb = (p.K + BK - 1) / BK;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int n_blk = t % p.tiles_n;
        const int m_blk = (t / p.tiles_n) % p.tiles_m;
        const int ks = t / (p.tiles_n * p.tiles_m);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(total_kb, kb0 + p.kb_per_split);
        const int m0 = m_blk * BM, n0 = n_blk * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          const int k0 = kb * BK;
          if (!A_MN) {
            tma_load_2d(sa, &tmA, &full_bar[stage], k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * 8192, &tmA, &full_bar[stage], m0 + j * 64, k0);
          }
          if (!B_MN) {
            tma_load_2d(sb, &tmB, &full_bar[stage], k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tmB, &full_bar[stage], n0 + j * 64, k0);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int ks = t / (p.tiles_n * p.tiles_m);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(total_kb, kb0 + p.kb_per_split);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint32_t sb = sa + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = A_MN ? make_smem_desc(sa + k * 2048, 8192, 1024) : make_smem_desc(sa + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc(sb + k * 2048, 8192, 1024) : make_smem_desc(sb + k * 32, 16, 1024);
            umma_bf16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs have read it
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int ew = warp - 2;
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int half = ew >> 2;             // which half of the BN columns
    constexpr int kColsPerWarp = BN / 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
      const int n_blk = t % p.tiles_n;
      const int m_blk = (t / p.tiles_n) % p.tiles_m;
      const int row = m_blk * BM + quad * 32 + lane;
      const int ncol0 = n_blk * BN + half * kColsPerWarp;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + acc * BN + half * kColsPerWarp + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < kColsPerWarp; c += 32) {
        uint32_t r[32];
        tmem_ld32(taddr0 + c, r);
        tmem_ld_wait();
        const int col0 = ncol0 + c;
        if (row < p.M && col0 < p.N) {
          const bool full = (col0 + 32 <= p.N);
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * p.alpha;
          if (p.bias) {
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) v[j] += __ldg(p.bias + col0 + j);
            }
          }
          if (p.out_pre) {
            __nv_bfloat16* o = p.out_pre + static_cast<size_t>(row) * p.ld_pre + col0;
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 u = make_uint4(pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]),
                                     pack_bf16x2(v[j + 4], v[j + 5]), pack_bf16x2(v[j + 6], v[j + 7]));
                *reinterpret_cast<uint4*>(o + j) = u;
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) o[j] = __float2bfloat16(v[j]);
            }
          }
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
          } else if (p.act == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
          }
          if (p.aux_mode) {
            const __nv_bfloat16* a = p.aux + static_cast<size_t>(row) * p.ld_aux + col0;
            float av[32];
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const uint4 u = __ldg(reinterpret_cast<const uint4*>(a + j));
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float2 f = __bfloat1622float2(h[q]);
                  av[j + 2 * q] = f.x; av[j + 2 * q + 1] = f.y;
                }
              }
            } else {
              for (int j = 0; j < 32; ++j) av[j] = (col0 + j < p.N) ? __bfloat162float(a[j]) : 0.0f;
            }
            if (p.aux_mode == 1) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= dgelu_erf(av[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = av[j] > 0.0f ? v[j] : 0.0f;
            }
          }
          if (p.resid) {
            const float* rp = p.resid + static_cast<size_t>(row) * p.ld_resid + col0;
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b = *reinterpret_cast<const float4*>(rp + j);
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) v[j] += rp[j];
            }
          }
          if (p.out_f32) {
            float* o = p.out_f32 + static_cast<size_t>(row) * p.ld_f32 + col0;
            if (p.atomic) {
              if (full) {  // 128-bit vector reductions (sm_90+): 4x fewer L2 atomic operations
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j), "f"(v[j]), "f"(v[j + 1]),
                               "f"(v[j + 2]), "f"(v[j + 3])
                               : "memory");
              } else {
                for (int j = 0; j < 32; ++j)
                  if (col0 + j < p.N) atomicAdd(o + j, v[j]);
              }
            } else if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) o[j] = v[j];
            }
          }
          if (p.out_bf16) {
            __nv_bfloat16* o = p.out_bf16 + static_cast<size_t>(row) * p.ld_bf16 + col0;
            if (full) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 u = make_uint4(pack_bf16x2(v[j], v[j + 1]), pack_bf16x2(v[j + 2], v[j + 3]),
                                     pack_bf16x2(v[j + 4], v[j + 5]), pack_bf16x2(v[j + 6], v[j + 7]));
                *reinterpret_cast<uint4*>(o + j) = u;
              }
            } else {
              for (int j = 0; j < 32; ++j)
                if (col0 + j < p.N) o[j] = __float2bfloat16(v[j]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const GemmArgs& a, const GemmDev& dev, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  CUtensorMap tmA, tmB;
  int rc;
  // K-major: global [rows, K] (K contiguous), box [rows_tile, 64].  MN-major: global [K, rows], box [64 k, 64 rows].
  if (!A_MN) rc = get_tmap_2d(a.A, a.M, a.K, a.lda, BM, BK, &tmA);
  else       rc = get_tmap_2d(a.A, a.K, a.M, a.lda, BK, 64, &tmA);
  if (rc) return rc;
  if (!B_MN) rc = get_tmap_2d(a.B, a.N, a.K, a.ldb, BN, BK, &tmB);
  else       rc = get_tmap_2d(a.B, a.K, a.N, a.ldb, BK, 64, &tmB);
  if (rc) return rc;
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    ETP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int tiles = dev.tiles_m * dev.tiles_n * dev.k_splits;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  if (g_prof_on) {
    char tag[160];
    snprintf(tag, sizeof(tag), "M%d N%d K%d bn%d %s%s ks%d%s%s%s%s%s%s", a.M, a.N, a.K, BN, A_MN ? "T" : "N", B_MN ? "T" : "N",
             dev.k_splits, a.bias ? " bias" : "", a.act ? (a.act == 1 ? " gelu" : " relu") : "", a.aux_mode ? " aux" : "",
             a.resid ? " resid" : "", a.out_f32 ? (a.atomic ? " red32" : " f32") : "", (a.out_bf16 ? " bf16" : ""));
    prof_tag(tag, 2.0 * a.M * a.N * a.K);
  }
  ETP_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(kThreads), Cfg::kSmemBytes, stream, tmA, tmB, dev));
  ETP_LAUNCHED();
  return ETP_OK;
}

int gemm(const GemmArgs& a, cudaStream_t stream) {
  ETP_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem");
  ETP_REQUIRE(a.A && a.B, "gemm: null operand");
  ETP_REQUIRE(a.out_f32 || a.out_bf16 || a.out_pre, "gemm: no output");
  ETP_REQUIRE(!a.atomic || (a.out_f32 && !a.out_bf16 && !a.out_pre && !a.resid), "gemm: atomic mode is fp32-out only");
  ETP_REQUIRE(a.k_splits >= 1 && (a.k_splits == 1 || a.atomic), "gemm: split-K needs atomic accumulation");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  ETP_REQUIRE(al16(a.bias) && al16(a.aux) && al16(a.resid) && al16(a.out_f32) && al16(a.out_bf16) && al16(a.out_pre),
              "gemm: epilogue pointers must be 16-byte aligned");
  ETP_REQUIRE((!a.resid || a.ld_resid % 4 == 0) && (!a.out_f32 || a.ld_f32 % 4 == 0) &&
                  (!a.out_bf16 || a.ld_bf16 % 8 == 0) && (!a.out_pre || a.ld_pre % 8 == 0) && (!a.aux_mode || a.ld_aux % 8 == 0),
              "gemm: epilogue leading dimensions must keep rows 16-byte aligned");
  GemmDev d;
  d.M = a.M; d.N = a.N; d.K = a.K;
  d.alpha = a.alpha;
  d.bias = a.bias; d.act = a.act; d.aux_mode = a.aux_mode; d.aux = a.aux; d.ld_aux = a.ld_aux;
  d.resid = a.resid; d.ld_resid = a.ld_resid; d.out_f32 = a.out_f32; d.ld_f32 = a.ld_f32; d.atomic = a.atomic;
  d.out_bf16 = a.out_bf16; d.ld_bf16 = a.ld_bf16; d.out_pre = a.out_pre; d.ld_pre = a.ld_pre;
  d.tiles_m = d.tiles_n = d.kb_per_split = 0; d.k_splits = a.k_splits;
  d.colsum = a.colsum;
  ETP_REQUIRE(!a.colsum || (a.k_splits == 1 && !a.atomic), "gemm: colsum needs whole-K tiles");
  // default: CTA-pair kernel (gemm_pair.cu).  ETP_GEMM_IMPL=1 selects the one-CTA kernel below (A/B measurements).
  static const bool legacy = [] { const char* e = getenv("ETP_GEMM_IMPL"); return e && e[0] == '1'; }();
  if (!legacy && a.N % 2 == 0) return gemm_pair(a, d, stream);  // (its epilogue moves column pairs)
  // tile-N choice: 256-wide tiles halve B re-reads from smem per flop; fall back to 128 when N is small
  // or when 256-wide tiles would leave most SMs idle.
  int bn = a.block_n;
  if (bn == 0) {
    const int tm = (a.M + BM - 1) / BM;
    const int t256 = tm * ((a.N + 255) / 256) * a.k_splits;
    bn = (a.N >= 256 && (a.N % 256 == 0 || a.N > 1024) && t256 >= num_sms() / 2) ? 256 : 128;
  }
  ETP_REQUIRE(bn == 128 || bn == 256, "gemm: block_n must be 128 or 256");
  d.tiles_m = (a.M + BM - 1) / BM;
  d.tiles_n = (a.N + bn - 1) / bn;
  d.k_splits = a.k_splits;
  const int total_kb = (a.K + BK - 1) / BK;
  d.kb_per_split = (total_kb + a.k_splits - 1) / a.k_splits;
  d.k_splits = (total_kb + d.kb_per_split - 1) / d.kb_per_split;  // drop empty splits
#define ETP_GEMM_DISPATCH(BN_)                                                            \
  if (!a.a_mn && !a.b_mn) return launch_gemm<BN_, false, false>(a, d, stream);            \
  if (!a.a_mn && a.b_mn) return launch_gemm<BN_, false, true>(a, d, stream);              \
  if (a.a_mn && a.b_mn) return launch_gemm<BN_, true, true>(a, d, stream);                \
  return launch_gemm<BN_, true, false>(a, d, stream);
  if (a.colsum) {  // the one-CTA kernel has no fused column sums: separate pass over the bf16 output
    ETP_REQUIRE(a.out_bf16 != nullptr, "gemm: colsum on the one-CTA kernel needs a bf16 output");
    GemmArgs b = a;
    b.colsum = nullptr;
    int rc = gemm(b, stream);
    if (rc) return rc;
    return colsum_bf16(a.out_bf16, a.M, a.N, a.ld_bf16, a.colsum, stream);
  }
  if (bn == 256) { ETP_GEMM_DISPATCH(256) }
  ETP_GEMM_DISPATCH(128)
#undef ETP_GEMM_DISPATCH
}

}  // namespace etp
