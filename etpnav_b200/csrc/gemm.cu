// bf16 GEMM on tcgen05 tensor cores (sm_100a), CTA pairs (cta_group::2).
//
//   D[M,N] = epilogue( alpha * sum_k A[m,k] * B[n,k] )
//
// Two CTAs of a cluster (one TPC: two SMs) own one 256 x BN output tile.  Each CTA stages ITS 128 rows of A
// and ITS half of the B tile (BN/2 rows) by TMA into its own shared memory; one thread of the leader CTA issues
// tcgen05.mma.cta_group::2 (M = 256, N = BN, K = 16), which reads A/B from both CTAs' shared memory and writes a
// 128 x BN fp32 accumulator into EACH CTA's TMEM.  Against the one-CTA 128 x BN tile this halves the B bytes
// every SM pulls through L2 and shared memory per flop (shared-memory operand reads drop from 96 to 64 B/clk/SM
// at BN = 256), which is what leaves shared-memory bandwidth for the epilogue.
//
// Pipelines (all mbarriers): full[s]  TMA of both CTAs -> leader's MMA thread   (lives in the leader)
//                            empty[s] MMA commit, multicast -> both CTAs' TMA producers
//                            tmem_full[a]  MMA commit, multicast -> both CTAs' epilogue warps
//                            tmem_empty[a] epilogue warps of both CTAs -> leader's MMA thread
// Two accumulator stages in TMEM overlap the epilogue of tile i with the main loop of tile i+1.
//
// Epilogue (8 warps per CTA): tcgen05.ld gives every thread one accumulator ROW (32 consecutive columns).
// Writing rows from that layout costs one 16-byte store per row per instruction (32 cache lines touched per
// warp instruction).  Instead each warp transposes its 32x32 chunk through a private padded shared-memory
// patch (STS.128 / LDS.32, conflict-free), after which lane = column: bias is one register, residual /
// activation-derivative loads and all stores are full 128-byte (fp32) or 64-byte (bf16) lines per warp
// instruction, and the residual of the next chunk is prefetched while the current one is processed.
//
// Each operand is either K-major (row-major [rows, K]) or MN-major ([K, rows], rows contiguous); the latter serves
// dgrad (B = W as stored) and wgrad (A = dY, B = X as stored) without transposes.  Split-K with fp32 vector
// reductions serves wgrad (few output tiles, long reduction).
// Replaces the cuBLAS calls behind torch.nn.Linear / torch.matmul on the reference path
// (vilmodel_cmt.py:108-110,326-328,151,178,190,654; common/transformer.py:174-181).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "gemm_dev.h"
#include "host.h"
#include "ops.h"

namespace etp {

namespace {

constexpr int BM = 128;  // accumulator rows per CTA (pair tile: 256)
constexpr int BK = 64;
// Epilogue warps per CTA (template parameter EW): 8 = 4 TMEM lane quadrants x 2 column halves (one more pipeline
// stage, 168 registers: next-chunk residual prefetch); 16 = 4 x 4 column quarters (four warps per scheduler hide the
// latency chains of the epilogue: TMEM load -> transpose -> math -> stores), 96 registers.
constexpr int kThreadsOf(int ew) { return 64 + ew * 32; }
constexpr int kTS = 36;  // staging pitch (floats): 16-byte aligned rows, conflict-free for STS.128 rows / LDS.32 columns
constexpr int kScratchBytesOf(int ew) { return ew * 32 * kTS * 4; }

template <int BN, int EW>
struct PairCfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = (BN / 2) * BK * 2;  // this CTA's half of the B tile
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (EW == 8) ? ((BN == 256) ? 5 : 7) : ((BN == 256) ? 4 : 6);
  static constexpr int kScratchBytes = kScratchBytesOf(EW);
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + kScratchBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// Grouped mode: several independent problems  dW_g += A_g^T.B_g  (the weight gradients of one transformer layer)
// share one persistent launch; a tile index maps to (problem, m block, n block).  Every tile owns its output and
// accumulates with a plain read-add-write.  One launch instead of seven removes six prologues / pipeline fills and
// lets the tiles of the small problems fill the waves of the large ones.
constexpr int kMaxGroup = 8;
struct GroupProblem {
  CUtensorMap tmA, tmB;
  int M, N, K;
  int tiles_n, tile_start;
  int ld;
  float* out;
};
struct GroupParams {
  int n, total_tiles;
  GroupProblem prob[kMaxGroup];
};

// what one tile index means (both modes)
struct TileRef {
  const CUtensorMap *tmA, *tmB;
  int m_blk, n_blk, kb0, kb1, g;
};

template <int BN, bool A_MN, bool B_MN, bool kGrouped, int EW>
ETP_DEVICE void gemm_body(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmDev& p, const GroupParams* gp) {
  using Cfg = PairCfg<BN, EW>;
  constexpr int kEpiWarps = EW;
  constexpr int kScratchBytes = Cfg::kScratchBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* scratch = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes + kScratchBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full = empty_bar + Cfg::kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    if constexpr (kGrouped) {
      for (int g = 0; g < gp->n; ++g) {
        tma_prefetch_desc(&gp->prob[g].tmA);
        tma_prefetch_desc(&gp->prob[g].tmB);
      }
    } else {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
    }
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 2);   // one arrival per CTA's producer (+ the bytes of both)
      mbar_init(&empty_bar[s], 1);  // the MMA commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 2 * kEpiWarps);  // epilogue warps of both CTAs
    }
    fence_barrier_init();
    fence_proxy_async();
  }
  __syncwarp();
  griddep_launch();  // PDL: the next kernel may start its own prologue
  griddep_wait();    // previous kernel complete; nothing above touched global memory or TMEM
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
  if (warp == 1) {
    tmem_alloc_pair(tmem_ptr, Cfg::kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // peer allocated too (execution barrier only) ...
  __syncthreads();     // ... and the TMEM base address written by warp 1 is visible to this CTA's warps
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int num_tiles = kGrouped ? gp->total_tiles : p.tiles_m * p.tiles_n * p.k_splits;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  auto decode = [&](int t) -> TileRef {
    TileRef r;
    if constexpr (kGrouped) {
      int g = 0;
      while (g + 1 < gp->n && t >= gp->prob[g + 1].tile_start) ++g;
      const GroupProblem& pr = gp->prob[g];
      const int local = t - pr.tile_start;
      r.g = g; r.tmA = &pr.tmA; r.tmB = &pr.tmB;
      r.n_blk = local % pr.tiles_n; r.m_blk = local / pr.tiles_n;
      r.kb0 = 0; r.kb1 = (pr.K + BK - 1) / BK;
    } else {
      const int total_kb = (p.K + BK - 1) / BK;
      r.g = 0; r.tmA = &tmA; r.tmB = &tmB;
      r.n_blk = t % p.tiles_n;
      r.m_blk = (t / p.tiles_n) % p.tiles_m;
      const int ks = t / (p.tiles_n * p.tiles_m);
      r.kb0 = ks * p.kb_per_split;
      r.kb1 = min(total_kb, r.kb0 + p.kb_per_split);
    }
    return r;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[0]), 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const TileRef tr = decode(t);
        const CUtensorMap* mA = tr.tmA;
        const CUtensorMap* mB = tr.tmB;
        const int m0 = tr.m_blk * (2 * BM) + static_cast<int>(rank) * BM;
        const int n0 = tr.n_blk * BN + static_cast<int>(rank) * (BN / 2);
        for (int kb = tr.kb0; kb < tr.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          uint8_t* sb = sa + Cfg::kABytes;
          const uint32_t bar = full_leader + 8 * stage;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
          else        mbar_arrive_cluster(bar);
          const int k0 = kb * BK;
          if (!A_MN) {
            tma_load_2d_pair(sa, mA, bar, k0, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d_pair(sa + j * 8192, mA, bar, m0 + j * 64, k0);
          }
          if (!B_MN) {
            tma_load_2d_pair(sb, mB, bar, k0, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 128; ++j) tma_load_2d_pair(sb + j * 8192, mB, bar, n0 + j * 64, k0);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA, one thread) =====================
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const TileRef tr = decode(t);
        const int kb0 = tr.kb0, kb1 = tr.kb1;
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
            umma_bf16_pair(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[stage]);  // frees the stage in both CTAs once these MMAs have read it
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit_pair(&tmem_full[acc]);  // accumulator complete -> epilogue warps of both CTAs
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs) =====================
    const int ew = warp - 2;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int half = ew >> 2;   // which slice (half / quarter) of the BN columns
    constexpr int kSlice = BN / (EW / 4);
    constexpr int kChunks = kSlice / 32;
    const uint32_t scr = smem_u32(scratch + ew * 32 * kTS);
    const uint32_t empty_leader = mapa_shared(smem_u32(&tmem_empty[0]), 0);
    // after the transpose a lane owns two adjacent columns (cp) of 16 rows (rh): every global access of a warp
    // instruction is two full 128-byte (fp32) / 64-byte (bf16) row segments
    const int cp = lane & 15;
    const int rh = lane >> 4;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      const TileRef tr = decode(t);
      const int n_blk = tr.n_blk, m_blk = tr.m_blk;
      GemmDev e = p;  // epilogue view of this tile's problem
      if constexpr (kGrouped) {
        const GroupProblem& pr = gp->prob[tr.g];
        e.M = pr.M; e.N = pr.N;
        e.out_f32 = pr.out; e.ld_f32 = pr.ld;
        e.resid = pr.out; e.ld_resid = pr.ld;  // dW += acc
      }
      const int row0 = m_blk * (2 * BM) + static_cast<int>(rank) * BM + quad * 32 + rh * 16;  // first row of this lane
      const int colbase = n_blk * BN + half * kSlice + 2 * cp;
      const int nrows = min(16, e.M - row0);  // <= 0: nothing of this lane's rows is inside the matrix

      // residual of chunk `c` in the transposed layout: 16 independent row segments in flight per lane
      float2 R[16];
      auto load_resid = [&](int c, float2 (&dst)[16]) {
        const int col = colbase + c * 32;
        // row pointers advance by the pitch (two adds per row) instead of being rebuilt from (row, col) per access
        const char* rp = reinterpret_cast<const char*>(e.resid + static_cast<size_t>(row0) * e.ld_resid + col);
        const int64_t pitch = static_cast<int64_t>(e.ld_resid) * 4;
        const bool ok = col < e.N;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          dst[k] = (ok && k < nrows) ? __ldg(reinterpret_cast<const float2*>(rp)) : make_float2(0.f, 0.f);
          rp += pitch;
        }
      };
      if (e.resid) load_resid(0, R);  // lands while the main loop of this tile is still running

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + acc * BN + half * kSlice + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c) {
        const int col = colbase + c * 32;
        const bool colok = col < e.N;
        uint32_t r[32];
        tmem_ld32(taddr0 + c * 32, r);
        float2 Rn[16];
        if constexpr (EW == 8) {
          if (e.resid && c + 1 < kChunks) load_resid(c + 1, Rn);  // next chunk's residual while this one is processed
        } else {
          if (e.resid && c > 0) load_resid(c, R);  // (chunk 0 was requested before the accumulator was ready)
        }
        uint32_t ax[16];
        if (e.aux_mode) {
          const char* ap = reinterpret_cast<const char*>(e.aux + static_cast<size_t>(row0) * e.ld_aux + col);
          const int64_t pitch = static_cast<int64_t>(e.ld_aux) * 2;
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            ax[k] = (colok && k < nrows) ? __ldg(reinterpret_cast<const unsigned int*>(ap)) : 0u;
            ap += pitch;
          }
        }
        tmem_ld_wait();
        if (c == kChunks - 1) {
          // accumulator stage drained: hand it back to the leader's MMA thread before the math and stores
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(empty_leader + 8 * acc);
        }
        // transpose the 32x32 chunk through this warp's padded patch: thread = row  ->  lane = (column pair, row half)
#pragma unroll
        for (int q = 0; q < 8; ++q)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(scr + (lane * kTS + 4 * q) * 4), "r"(r[4 * q]),
                       "r"(r[4 * q + 1]), "r"(r[4 * q + 2]), "r"(r[4 * q + 3])
                       : "memory");
        __syncwarp();
        float v[32];
#pragma unroll
        for (int k = 0; k < 16; ++k)
          asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v[2 * k]), "=f"(v[2 * k + 1])
                       : "r"(scr + ((rh * 16 + k) * kTS + 2 * cp) * 4)
                       : "memory");
        __syncwarp();

        if (nrows > 0 && colok) {
          float2 bias = make_float2(0.f, 0.f);
          if (e.bias) bias = __ldg(reinterpret_cast<const float2*>(e.bias + col));
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            v[2 * k] = fmaf(v[2 * k], e.alpha, bias.x);
            v[2 * k + 1] = fmaf(v[2 * k + 1], e.alpha, bias.y);
          }
          if (e.out_pre && !e.pre_mode) {
            char* o = reinterpret_cast<char*>(e.out_pre + static_cast<size_t>(row0) * e.ld_pre + col);
            const int64_t pitch = static_cast<int64_t>(e.ld_pre) * 2;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              st_global_b32_if(o, pack_bf16x2(v[2 * k], v[2 * k + 1]), k < nrows);
              o += pitch;
            }
          }
          if (e.act == 1 && e.pre_mode) {
            // GELU and its derivative from one Phi / exp evaluation; the derivative is what the backward multiplies by
            char* o = reinterpret_cast<char*>(e.out_pre + static_cast<size_t>(row0) * e.ld_pre + col);
            const int64_t pitch = static_cast<int64_t>(e.ld_pre) * 2;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              float c0, e0, c1, e1;
              phi_parts(v[2 * k], c0, e0);
              phi_parts(v[2 * k + 1], c1, e1);
              float d0 = fmaf(v[2 * k] * 0.39894228040143267794f, e0, c0);
              float d1 = fmaf(v[2 * k + 1] * 0.39894228040143267794f, e1, c1);
              if (e.drop_thr) {  // d/dpre of dropout(gelu(pre)) = mask/(1-p) * gelu'(pre)
                float m0, m1;
                drop_mul2(Drop{e.drop_key, e.drop_thr, e.drop_scale},
                          static_cast<uint32_t>(row0 + k) * static_cast<uint32_t>(e.N) + static_cast<uint32_t>(col), m0, m1);
                d0 *= m0; d1 *= m1;
              }
              st_global_b32_if(o, pack_bf16x2(d0, d1), k < nrows);
              o += pitch;
              v[2 * k] *= c0;
              v[2 * k + 1] *= c1;
            }
          } else if (e.act == 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_fast(v[i]);
          } else if (e.act == 2) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.0f);
          }
          if (e.aux_mode == 1) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              v[2 * k] *= dgelu_fast(__uint_as_float(ax[k] << 16));
              v[2 * k + 1] *= dgelu_fast(__uint_as_float(ax[k] & 0xffff0000u));
            }
          } else if (e.aux_mode == 2) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              v[2 * k] = (__uint_as_float(ax[k] << 16) > 0.0f) ? v[2 * k] : 0.0f;
              v[2 * k + 1] = (__uint_as_float(ax[k] & 0xffff0000u) > 0.0f) ? v[2 * k + 1] : 0.0f;
            }
          }
          if (e.aux_mode == 3) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              v[2 * k] *= __uint_as_float(ax[k] << 16);
              v[2 * k + 1] *= __uint_as_float(ax[k] & 0xffff0000u);
            }
          }
          if (e.drop_thr) {
            const Drop dr{e.drop_key, e.drop_thr, e.drop_scale};
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              float m0, m1;
              drop_mul2(dr, static_cast<uint32_t>(row0 + k) * static_cast<uint32_t>(e.N) + static_cast<uint32_t>(col), m0, m1);
              v[2 * k] *= m0;
              v[2 * k + 1] *= m1;
            }
          }
          if (e.resid) {
#pragma unroll
            for (int k = 0; k < 16; ++k) { v[2 * k] += R[k].x; v[2 * k + 1] += R[k].y; }
          }
          if (e.out_f32) {
            float* o = e.out_f32 + static_cast<size_t>(row0) * e.ld_f32 + col;
            if (e.atomic) {
#pragma unroll
              for (int k = 0; k < 16; ++k)
                if (k < nrows)
                  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(o + static_cast<size_t>(k) * e.ld_f32), "f"(v[2 * k]),
                               "f"(v[2 * k + 1])
                               : "memory");
            } else {
              char* ob = reinterpret_cast<char*>(o);
              const int64_t pitch = static_cast<int64_t>(e.ld_f32) * 4;
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                st_global_v2f32_if(ob, v[2 * k], v[2 * k + 1], k < nrows);
                ob += pitch;
              }
            }
          }
          if (e.out_bf16) {
            char* o = reinterpret_cast<char*>(e.out_bf16 + static_cast<size_t>(row0) * e.ld_bf16 + col);
            const int64_t pitch = static_cast<int64_t>(e.ld_bf16) * 2;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              st_global_b32_if(o, pack_bf16x2(v[2 * k], v[2 * k + 1]), k < nrows);
              o += pitch;
            }
          }
        }
        if constexpr (EW == 8) {
          if (e.resid && c + 1 < kChunks) {
#pragma unroll
            for (int k = 0; k < 16; ++k) R[k] = Rn[k];
          }
        }
        if (e.colsum) {
          // bias gradient: column sums of the final values (rows outside the matrix contribute 0)
          float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            s0 += (k < nrows && colok) ? v[2 * k] : 0.0f;
            s1 += (k < nrows && colok) ? v[2 * k + 1] : 0.0f;
          }
          s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
          s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
          if (rh == 0 && colok)
            asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(e.colsum + col), "f"(s0), "f"(s1) : "memory");
        }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  // both CTAs stay resident until the pair is done: the leader's MMAs read the peer's shared memory and the
  // peer's warps arrive on the leader's barriers
  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  }
}

template <int BN, bool A_MN, bool B_MN, int EW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsOf(EW), 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmDev p) {
  gemm_body<BN, A_MN, B_MN, false, EW>(tmA, tmB, p, nullptr);
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsOf(8), 1)
gemm_tcgen05_grouped_kernel(const __grid_constant__ GroupParams gp, const GemmDev p) {
  gemm_body<BN, A_MN, B_MN, true, 8>(gp.prob[0].tmA, gp.prob[0].tmB, p, &gp);
}

template <int BN, bool A_MN, bool B_MN, int EW>
int launch_pair_ew(const GemmArgs& a, const GemmDev& dev, cudaStream_t stream) {
  using Cfg = PairCfg<BN, EW>;
  CUtensorMap tmA, tmB;
  int rc;
  // K-major: global [rows, K] (K contiguous), box [rows_per_cta, 64].  MN-major: global [K, rows], box [64 k, 64 rows].
  if (!A_MN) rc = get_tmap_2d(a.A, a.M, a.K, a.lda, BM, BK, &tmA);
  else       rc = get_tmap_2d(a.A, a.K, a.M, a.lda, BK, 64, &tmA);
  if (rc) return rc;
  if (!B_MN) rc = get_tmap_2d(a.B, a.N, a.K, a.ldb, BN / 2, BK, &tmB);
  else       rc = get_tmap_2d(a.B, a.K, a.N, a.ldb, BK, 64, &tmB);
  if (rc) return rc;
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN, EW>;
  static bool attr_set = false;
  if (!attr_set) {
    ETP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int tiles = dev.tiles_m * dev.tiles_n * dev.k_splits;
  const int pairs = num_sms() / 2;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  if (g_prof_on) {
    char tag[160];
    snprintf(tag, sizeof(tag), "M%d N%d K%d bn%d ew%d %s%s ks%d%s%s%s%s%s%s", a.M, a.N, a.K, BN, EW, A_MN ? "T" : "N",
             B_MN ? "T" : "N", dev.k_splits, a.bias ? " bias" : "", a.act ? (a.act == 1 ? " gelu" : " relu") : "",
             a.aux_mode ? " aux" : "", a.resid ? " resid" : "", a.out_f32 ? (a.atomic ? " red32" : " f32") : "",
             (a.out_bf16 ? " bf16" : ""));
    prof_tag(tag, 2.0 * a.M * a.N * a.K);
  }
  ETP_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(kThreadsOf(EW)), Cfg::kSmemBytes, stream, tmA, tmB, dev));
  ETP_LAUNCHED();
  return ETP_OK;
}

template <int BN, bool A_MN, bool B_MN>
int launch_pair(const GemmArgs& a, const GemmDev& dev, cudaStream_t stream) {
  // 16 epilogue warps for every single-problem launch: measured on one box (tests/run_gpu_ab.sh) the c3 training step
  // runs 5.67 ms with 16 everywhere, 5.85 ms with 16 only for GELU / derivative / dropout epilogues, 5.96 ms with 8.
  // The grouped weight-gradient launch (main-loop bound, 80 k-blocks per tile) keeps 8 warps and 5 stages.
  // ETP_GEMM_EW=8 forces the 8-warp variant (A/B measurements).
  bool heavy = true;
  static const int forced = [] { const char* e = getenv("ETP_GEMM_EW"); return e ? atoi(e) : 0; }();
  if (forced == 8) heavy = false;
  return heavy ? launch_pair_ew<BN, A_MN, B_MN, 16>(a, dev, stream) : launch_pair_ew<BN, A_MN, B_MN, 8>(a, dev, stream);
}

template <int BN>
int launch_grouped_tt(const GemmArgs* a, int n, cudaStream_t stream) {
  using Cfg = PairCfg<BN, 8>;
  GroupParams gp;
  memset(&gp, 0, sizeof(gp));
  gp.n = n;
  int tiles = 0;
  double flops = 0.0;
  for (int g = 0; g < n; ++g) {
    GroupProblem& pr = gp.prob[g];
    int rc = get_tmap_2d(a[g].A, a[g].K, a[g].M, a[g].lda, BK, 64, &pr.tmA);  // MN-major: global [K, M], box [64 k, 64 m]
    if (rc) return rc;
    rc = get_tmap_2d(a[g].B, a[g].K, a[g].N, a[g].ldb, BK, 64, &pr.tmB);
    if (rc) return rc;
    pr.M = a[g].M; pr.N = a[g].N; pr.K = a[g].K;
    pr.tiles_n = (a[g].N + BN - 1) / BN;
    pr.tile_start = tiles;
    pr.ld = a[g].ld_f32;
    pr.out = a[g].out_f32;
    tiles += ((a[g].M + 2 * BM - 1) / (2 * BM)) * pr.tiles_n;
    flops += 2.0 * a[g].M * a[g].N * a[g].K;
  }
  gp.total_tiles = tiles;
  GemmDev d;
  memset(&d, 0, sizeof(d));
  d.alpha = 1.0f;
  auto kern = gemm_tcgen05_grouped_kernel<BN, true, true>;
  static bool attr_set = false;
  if (!attr_set) {
    ETP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  const int pairs = num_sms() / 2;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  if (g_prof_on) {
    char tag[96];
    snprintf(tag, sizeof(tag), "grouped wgrad x%d bn%d TT K%d tiles%d", n, BN, a[0].K, tiles);
    prof_tag(tag, flops);
  }
  ETP_CHECK_CUDA(launch_pdl(kern, dim3(grid), dim3(kThreadsOf(8)), Cfg::kSmemBytes, stream, gp, d));
  ETP_LAUNCHED();
  return ETP_OK;
}

}  // namespace

int gemm(const GemmArgs& a, cudaStream_t stream) {
  ETP_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem");
  ETP_REQUIRE(a.A && a.B, "gemm: null operand");
  ETP_REQUIRE(a.out_f32 || a.out_bf16 || a.out_pre, "gemm: no output");
  ETP_REQUIRE(a.N % 2 == 0, "gemm: N must be even (the epilogue moves column pairs)");
  ETP_REQUIRE(!a.atomic || (a.out_f32 && !a.out_bf16 && !a.out_pre && !a.resid), "gemm: atomic mode is fp32-out only");
  ETP_REQUIRE(a.k_splits >= 1 && (a.k_splits == 1 || a.atomic), "gemm: split-K needs atomic accumulation");
  ETP_REQUIRE(!a.colsum || (a.k_splits == 1 && !a.atomic), "gemm: colsum needs whole-K tiles");
  ETP_REQUIRE(!a.pre_mode || (a.act == 1 && a.out_pre), "gemm: pre_mode 1 stores gelu'(pre) and needs act = gelu");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  ETP_REQUIRE(al16(a.bias) && al16(a.aux) && al16(a.resid) && al16(a.out_f32) && al16(a.out_bf16) && al16(a.out_pre) &&
                  al16(a.colsum),
              "gemm: epilogue pointers must be 16-byte aligned");
  ETP_REQUIRE((!a.resid || a.ld_resid % 4 == 0) && (!a.out_f32 || a.ld_f32 % 4 == 0) &&
                  (!a.out_bf16 || a.ld_bf16 % 8 == 0) && (!a.out_pre || a.ld_pre % 8 == 0) && (!a.aux_mode || a.ld_aux % 8 == 0),
              "gemm: epilogue leading dimensions must keep rows 16-byte aligned");
  GemmDev d;
  d.M = a.M; d.N = a.N; d.K = a.K;
  d.alpha = a.alpha;
  d.bias = a.bias; d.act = a.act; d.aux_mode = a.aux_mode; d.aux = a.aux; d.ld_aux = a.ld_aux;
  d.resid = a.resid; d.ld_resid = a.ld_resid; d.out_f32 = a.out_f32; d.ld_f32 = a.ld_f32; d.atomic = a.atomic;
  d.out_bf16 = a.out_bf16; d.ld_bf16 = a.ld_bf16; d.out_pre = a.out_pre; d.ld_pre = a.ld_pre; d.pre_mode = a.pre_mode;
  d.colsum = a.colsum;
  d.drop_key = a.drop_key; d.drop_thr = a.drop_thr; d.drop_scale = a.drop_scale;
  ETP_REQUIRE(!a.drop_thr || (a.k_splits == 1 && !a.atomic && static_cast<int64_t>(a.M) * a.N < (int64_t(1) << 32)),
              "gemm: dropout needs whole-K tiles and < 2^32 elements");
  d.k_splits = a.k_splits;
  // tile-N: 256-wide pair tiles unless N is small / not a multiple of 256, or they would leave most pairs idle
  int bn = a.block_n;
  const int tm = (a.M + 2 * BM - 1) / (2 * BM);
  if (bn == 0) {
    // minimise  waves x (k-blocks + epilogue) in units of one 256x256x64 MMA block (see wgrad() in planner_bwd.cu)
    const int pairs = num_sms() / 2;
    const int kbs = ((a.K + BK - 1) / BK + a.k_splits - 1) / a.k_splits;
    double best = 1e30;
    for (int cand = 128; cand <= 256; cand += 128) {
      if (cand == 256 && a.N < 256) continue;
      const int tiles = tm * ((a.N + cand - 1) / cand) * a.k_splits;
      const int waves = (tiles + pairs - 1) / pairs;
      const double w = cand / 256.0;
      const double cost = waves * (kbs * w + 8.0 * w + 4.0);
      if (cost < best) { best = cost; bn = cand; }
    }
  }
  ETP_REQUIRE(bn == 128 || bn == 256, "gemm: block_n must be 128 or 256");
  d.tiles_m = tm;
  d.tiles_n = (a.N + bn - 1) / bn;
  const int total_kb = (a.K + BK - 1) / BK;
  d.kb_per_split = (total_kb + a.k_splits - 1) / a.k_splits;
  d.k_splits = (total_kb + d.kb_per_split - 1) / d.kb_per_split;  // drop empty splits
#define ETP_PAIR_DISPATCH(BN_)                                                          \
  if (!a.a_mn && !a.b_mn) return launch_pair<BN_, false, false>(a, d, stream);          \
  if (!a.a_mn && a.b_mn) return launch_pair<BN_, false, true>(a, d, stream);            \
  if (a.a_mn && a.b_mn) return launch_pair<BN_, true, true>(a, d, stream);              \
  return launch_pair<BN_, true, false>(a, d, stream);
  if (bn == 256) { ETP_PAIR_DISPATCH(256) }
  ETP_PAIR_DISPATCH(128)
#undef ETP_PAIR_DISPATCH
}

// Several weight-gradient problems  out_g[M_g, N_g] += A_g^T . B_g  (A_g stored [K_g, M_g], B_g stored [K_g, N_g]) in ONE
// persistent launch.  Every problem: a_mn = b_mn = 1, fp32 output accumulated in place, no other epilogue.
int gemm_grouped_wgrad(const GemmArgs* a, int n, cudaStream_t stream) {
  ETP_REQUIRE(a != nullptr && n >= 1 && n <= kMaxGroup, "gemm_grouped_wgrad: 1..8 problems");
  bool all256 = true;
  for (int g = 0; g < n; ++g) {
    ETP_REQUIRE(a[g].M > 0 && a[g].N > 0 && a[g].K > 0 && a[g].A && a[g].B && a[g].out_f32, "gemm_grouped_wgrad: bad problem");
    ETP_REQUIRE(a[g].a_mn && a[g].b_mn && a[g].N % 2 == 0 && a[g].ld_f32 % 4 == 0 &&
                    (reinterpret_cast<uintptr_t>(a[g].out_f32) & 15) == 0,
                "gemm_grouped_wgrad: operands must be MN-major, outputs 16-byte aligned");
    all256 = all256 && (a[g].N % 256 == 0);
  }
  // tile width by the same cost model as the single-problem path: waves x (k-blocks + epilogue)
  const int pairs = num_sms() / 2;
  double best = 1e30;
  int bn = 128;
  for (int cand = 128; cand <= 256; cand += 128) {
    if (cand == 256 && !all256) continue;
    int tiles = 0, kbmax = 0;
    for (int g = 0; g < n; ++g) {
      tiles += ((a[g].M + 2 * BM - 1) / (2 * BM)) * ((a[g].N + cand - 1) / cand);
      kbmax = max(kbmax, (a[g].K + BK - 1) / BK);
    }
    const int waves = (tiles + pairs - 1) / pairs;
    const double w = cand / 256.0;
    const double cost = waves * (kbmax * w + 8.0 * w + 4.0);
    if (cost < best) { best = cost; bn = cand; }
  }
  return bn == 256 ? launch_grouped_tt<256>(a, n, stream) : launch_grouped_tt<128>(a, n, stream);
}

}  // namespace etp
