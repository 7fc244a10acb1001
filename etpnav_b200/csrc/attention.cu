// Multi-head attention, head dim 64, CUDA-core variant (used for the panorama encoder where the
// sequence is <= 16 views, and as the always-available baseline of the tensor-core kernel in
// attention_tc.cu).  softmax(scale * q.k^T + bias) . v with
//   bias[b,q,k] = (key_valid[b,k] ? 0 : mask_value) + pair_w * pair[b,q,k] + pair_b
// which covers BertOutAttention (vilmodel_cmt.py:325-352: text mask -10000), BertSelfAttention
// (:103-141: node mask -10000 + sprel bias, :391-393) and nn.MultiheadAttention's key_padding_mask
// (common/transformer.py:176: -inf).  One CTA = 16 query rows of one (batch, head); K/V of the head are
// staged once in shared memory (bf16), each warp owns a query row at a time with keys across lanes.
#include "common.cuh"
#include "host.h"
#include "ops.h"

namespace etp {

constexpr int kD = 64;
constexpr int kKStride = 66;  // bf16 elements per K row in smem (33 words: conflict-free across rows)
constexpr int kQTile = 16;

__global__ void __launch_bounds__(256) attention_fwd_kernel(const AttnArgs a, int sk_pad) {
  griddep_launch();  // PDL (common.cuh): let the next kernel get scheduled ...
  griddep_wait();    // ... and wait for the previous one before touching memory
  extern __shared__ uint8_t smem[];
  bf16* Ks = reinterpret_cast<bf16*>(smem);
  bf16* Vs = Ks + static_cast<size_t>(sk_pad) * kKStride;
  float* qs = reinterpret_cast<float*>(Vs + static_cast<size_t>(sk_pad) * kD);
  float* ps = qs + 8 * kD;
  float* kb = ps + 8 * sk_pad;  // per-key bias (mask), [sk_pad]

  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * kQTile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int Sk = a.Sk;

  // stage K, V (bf16) and the per-key mask
  const int kvb = a.kv_rows ? __ldg(a.kv_rows + b) : b;  // row of the K / V tensor this batch element reads
  const bf16* kg = a.k + static_cast<size_t>(kvb) * Sk * a.ldk + h * kD;
  const bf16* vg = a.v + static_cast<size_t>(kvb) * Sk * a.ldv + h * kD;
  for (int i = threadIdx.x; i < Sk * 8; i += 256) {
    const int r = i >> 3, c = (i & 7) * 8;
    const uint4 kk = *reinterpret_cast<const uint4*>(kg + static_cast<size_t>(r) * a.ldk + c);
    uint32_t* kd = reinterpret_cast<uint32_t*>(Ks + r * kKStride + c);
    kd[0] = kk.x; kd[1] = kk.y; kd[2] = kk.z; kd[3] = kk.w;
    *reinterpret_cast<uint4*>(Vs + r * kD + c) = *reinterpret_cast<const uint4*>(vg + static_cast<size_t>(r) * a.ldv + c);
  }
  for (int j = threadIdx.x; j < sk_pad; j += 256) {
    float m = -INFINITY;
    if (j < Sk) m = (a.key_valid == nullptr || a.key_valid[static_cast<size_t>(b) * Sk + j]) ? 0.f : a.mask_value;
    kb[j] = m;
  }
  __syncthreads();

  const float pair_w = a.pair_w_dev ? __ldg(a.pair_w_dev) : a.pair_w;
  const float pair_b = a.pair_b_dev ? __ldg(a.pair_b_dev) : a.pair_b;
  float* myq = qs + warp * kD;
  float* myp = ps + warp * sk_pad;
  const int nk = sk_pad / 32;
  for (int qi = warp; qi < kQTile; qi += 8) {
    const int q = q0 + qi;
    if (q >= a.Sq) break;
    const bf16* qg = a.q + (static_cast<size_t>(b) * a.Sq + q) * a.ldq + h * kD;
    {
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(qg + lane * 2));
      myq[lane * 2] = f.x * a.scale;
      myq[lane * 2 + 1] = f.y * a.scale;
    }
    __syncwarp();
    const float* pair = a.pair ? a.pair + (static_cast<size_t>(b) * a.Sq + q) * Sk : nullptr;
    float mx = -INFINITY;
    for (int t = 0; t < nk; ++t) {
      const int j = t * 32 + lane;
      float s = -INFINITY;
      if (j < Sk) {
        const uint32_t* kr = reinterpret_cast<const uint32_t*>(Ks + j * kKStride);
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < kD / 2; ++d) {
          const uint32_t u = kr[d];
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
          acc += myq[2 * d] * f.x + myq[2 * d + 1] * f.y;
        }
        s = acc + kb[j];
        if (pair) s += pair_w * pair[j] + pair_b;
      }
      myp[j] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int t = 0; t < nk; ++t) {
      const int j = t * 32 + lane;
      const float e = __expf(myp[j] - mx);  // exp(-inf) = 0 for padded / -inf-masked keys
      myp[j] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    if (a.drop_thr) {  // dropout on the normalised probabilities: the row sum above stays that of the undropped row
      const Drop dr{a.drop_key, a.drop_thr, a.drop_scale};
      const uint32_t e0 = static_cast<uint32_t>((static_cast<size_t>(b) * a.heads + h) * a.Sq + q) * static_cast<uint32_t>(Sk);
      for (int t = 0; t < nk; ++t) {
        const int j = t * 32 + lane;
        if (j < Sk) myp[j] *= drop_mul(dr, e0 + j);
      }
    }
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < Sk; ++j) {
      const float pj = myp[j];
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(Vs + j * kD + lane * 2));
      o0 += pj * f.x;
      o1 += pj * f.y;
    }
    const float inv = 1.0f / sum;
    bf16* og = a.out + (static_cast<size_t>(b) * a.Sq + q) * a.ldo + h * kD;
    *reinterpret_cast<uint32_t*>(og + lane * 2) = pack_bf16x2(o0 * inv, o1 * inv);
    if (a.lse && lane == 0) a.lse[(static_cast<size_t>(b) * a.heads + h) * a.Sq + q] = mx + __logf(sum);
    __syncwarp();
  }
}

int attention_fwd(const AttnArgs& a, cudaStream_t stream) {
  ETP_REQUIRE(a.B > 0 && a.Sq > 0 && a.Sk > 0 && a.heads > 0, "attention: empty problem");
  ETP_REQUIRE(a.q && a.k && a.v && a.out, "attention: null argument");
  ETP_REQUIRE(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 2 == 0, "attention: pitches must be 16-byte multiples");
  ETP_REQUIRE(a.Sk <= 1024, "attention: Sk > 1024 not supported");
  const int sk_pad = (a.Sk + 31) / 32 * 32;
  const size_t smem = static_cast<size_t>(sk_pad) * kKStride * 2 + static_cast<size_t>(sk_pad) * kD * 2 + 8 * kD * 4 +
                      8 * static_cast<size_t>(sk_pad) * 4 + static_cast<size_t>(sk_pad) * 4;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    ETP_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    configured = 220 * 1024;
  }
  ETP_REQUIRE(smem <= 220 * 1024, "attention: K/V do not fit shared memory");
  dim3 grid((a.Sq + kQTile - 1) / kQTile, a.heads, a.B);
  ETP_CHECK_CUDA(launch_pdl(attention_fwd_kernel, dim3(grid), dim3(256), smem, stream, a, sk_pad));
  ETP_LAUNCHED();
  return ETP_OK;
}

int attention_tc_fwd(const AttnArgs& a, cudaStream_t stream);  // attention_tc.cu
bool attention_tc_supported(const AttnArgs& a);

int attention_dispatch(const AttnArgs& a, cudaStream_t stream) {
  const bool tiny = a.Sk < 32 && a.Sq < 32;  // e.g. panorama views: not worth a 128-row tensor-core tile
  if (!tiny && attention_tc_supported(a)) return attention_tc_fwd(a, stream);
  return attention_fwd(a, stream);
}

}  // namespace etp
