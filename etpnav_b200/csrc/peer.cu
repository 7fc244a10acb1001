// Data-parallel update over NVLink peer memory: reduce-scatter of the gradients, AdamW and the all-gather of the new
// parameters in ONE kernel (the exchange step of the reference's DistributedDataParallel + torch.optim.AdamW,
// vlnce_baselines/ss_trainer_ETP.py:211-213; NCCL all-reduce + a replicated update is the library-call baseline of the
// same step, PlannerTrainer grad_comm="fp32"/"bf16").
//
// Every rank maps the flat gradient / parameter / bf16-image buffers of all ranks (CUDA IPC, one process per GPU;
// NVSwitch gives every pair the full link rate).  A trainable run [x, y) of a gradient bucket is cut into `world` equal
// sub-slices; rank r OWNS sub-slice r:
//     g      = sum over ranks (fixed order 0..world-1) of grad_rank[i]          peer LOADS  (reduce-scatter)
//     p,m,v  = AdamW(p, g / world, m, v)        m, v exist on the owner only    1/world of the optimizer state and work
//     param_rank[i] = p, image_rank[i] = bf16(p)  for every rank                peer STORES (all-gather)
// so one element's update is computed once, from one summation order, and every rank receives the same bits.
// Ordering between ranks is by flags in peer memory, one 32-bit word per (kind, bucket, rank) holding the step number:
//   READY(b): rank r finished the backward kernels that write bucket b's gradients (and read its weights) in this step;
//   DONE(b):  rank r finished reading bucket b's gradients from, and writing bucket b's new parameters to, every rank.
// Per bucket, on the update stream:  signal READY -> wait READY of all ranks -> fused kernel(s) -> signal DONE;
// the compute stream waits for DONE of all buckets and ranks before the next step zeroes the gradients / reads weights.
// Signals are their own tiny launches (a kernel boundary orders them after the data they publish; they are not launched
// with the programmatic-serialization attribute), waits spin on the rank's OWN flag block with a timeout that raises an
// error word instead of hanging the device.
#include <cudaTypedefs.h>
#include <string.h>

#include "common.cuh"
#include "host.h"
#include "../../include/etpnav_b200.h"

namespace etp {

using bf16 = __nv_bfloat16;
constexpr int kMaxRanks = 8;
constexpr int kMaxBuckets = 32;

struct PeerDev {
  int world, rank;
  float* grad[kMaxRanks];
  float* param[kMaxRanks];
  bf16* image[kMaxRanks];
  uint32_t* flags[kMaxRanks];
};

__device__ int g_peer_error = 0;

ETP_DEVICE void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
ETP_DEVICE uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
ETP_DEVICE float4 ld_peer_f4(const float4* p) {   // L1-bypassing 16-byte load (the line lives in another GPU's memory)
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
ETP_DEVICE unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void peer_signal_kernel(PeerDev g, int kind, int bucket, uint32_t value) {
  __threadfence_system();
  if (static_cast<int>(threadIdx.x) < g.world)
    st_release_sys(g.flags[threadIdx.x] + (kind * kMaxBuckets + bucket) * kMaxRanks + g.rank, value);
}

__global__ void peer_wait_kernel(PeerDev g, int kind, int b_lo, int b_hi, uint32_t value, unsigned long long timeout_ns) {
  const int cnt = (b_hi - b_lo) * g.world;
  const unsigned long long t0 = global_ns();
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int b = b_lo + i / g.world, r = i % g.world;
    const uint32_t* p = g.flags[g.rank] + (kind * kMaxBuckets + b) * kMaxRanks + r;
    while (static_cast<int32_t>(ld_acquire_sys(p) - value) < 0) {
      if (*reinterpret_cast<volatile int*>(&g_peer_error) != 0) break;   // an earlier wait already gave up: do not stack timeouts
      if (global_ns() - t0 > timeout_ns) {
        atomicExch(&g_peer_error, 1 + kind);
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
  __threadfence_system();
}

// Grid-stride over the owner's n4 float4 of one sub-slice that starts at float4 index off4 of the flat buffers.  A peer load
// takes ~2 us over NVLink, so the link is only full with megabytes in flight: every thread issues kU x world 16-byte loads
// (kU x world = 16) before it consumes the first one — 64 KB per CTA.
template <int kWorld, int kU>
__global__ void __launch_bounds__(256) peer_reduce_adamw_kernel(PeerDev g, int64_t off4, int64_t n4, float4* __restrict__ m,
                                                                float4* __restrict__ v, float lr, float b1, float b2, float eps,
                                                                float decay, float bc1, float bc2_sqrt, float gscale,
                                                                int write_reduced) {
  const int me = g.rank;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i0 < n4; i0 += stride * kU) {
    float4 part[kU][kWorld];
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < n4) {
#pragma unroll
        for (int r = 0; r < kWorld; ++r)
          if (r < g.world) part[u][r] = ld_peer_f4(reinterpret_cast<const float4*>(g.grad[r]) + off4 + i);
      }
    }
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= n4) break;
      const int64_t j = off4 + i;
      float4 pv = reinterpret_cast<const float4*>(g.param[me])[j];
      float4 mv = m[i], vv = v[i];
      float4 gs = part[u][0];
#pragma unroll
      for (int r = 1; r < kWorld; ++r)
        if (r < g.world) { gs.x += part[u][r].x; gs.y += part[u][r].y; gs.z += part[u][r].z; gs.w += part[u][r].w; }
      if (write_reduced) reinterpret_cast<float4*>(g.grad[me])[j] = gs;   // own sub-slice: nobody else reads it
      float* pp = reinterpret_cast<float*>(&pv);
      const float* gg = reinterpret_cast<const float*>(&gs);
      float* mm = reinterpret_cast<float*>(&mv);
      float* vq = reinterpret_cast<float*>(&vv);
#pragma unroll
      for (int k = 0; k < 4; ++k) {     // the arithmetic of adamw_kernel (pack_bwd.cu), operation for operation
        const float gr = gg[k] * gscale;
        pp[k] *= decay;
        mm[k] = b1 * mm[k] + (1.0f - b1) * gr;
        vq[k] = b2 * vq[k] + (1.0f - b2) * gr * gr;
        const float denom = sqrtf(vq[k]) / bc2_sqrt + eps;
        pp[k] -= (lr / bc1) * (mm[k] / denom);
      }
      m[i] = mv;
      v[i] = vv;
      const uint2 img = make_uint2(pack_bf16x2(pp[0], pp[1]), pack_bf16x2(pp[2], pp[3]));
#pragma unroll
      for (int r = 0; r < kWorld; ++r)
        if (r < g.world) {
          reinterpret_cast<float4*>(g.param[r])[j] = pv;
          reinterpret_cast<uint2*>(g.image[r])[j] = img;
        }
    }
  }
}

static int to_dev(const etp_peer_group* g, PeerDev* d) {
  ETP_REQUIRE(g != nullptr, "peer group: null");
  ETP_REQUIRE(g->world >= 1 && g->world <= kMaxRanks && g->rank >= 0 && g->rank < g->world, "peer group: bad world / rank");
  d->world = g->world;
  d->rank = g->rank;
  for (int r = 0; r < kMaxRanks; ++r) {
    d->grad[r] = r < g->world ? g->grad[r] : nullptr;
    d->param[r] = r < g->world ? g->param[r] : nullptr;
    d->image[r] = r < g->world ? reinterpret_cast<bf16*>(g->image[r]) : nullptr;
    d->flags[r] = r < g->world ? g->flags[r] : nullptr;
    if (r < g->world) ETP_REQUIRE(d->grad[r] && d->param[r] && d->image[r] && d->flags[r], "peer group: null buffer");
  }
  return ETP_OK;
}

static PFN_cuMemGetAddressRange_v3020 g_get_range = nullptr;

}  // namespace etp

using namespace etp;
#define ETP_API __attribute__((visibility("default")))
static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

ETP_API int etp_ipc_export(const void* ptr, void* handle64, int64_t* offset) {
  ETP_REQUIRE(ptr && handle64 && offset, "etp_ipc_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (g_get_range == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    ETP_CHECK_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q));
    ETP_REQUIRE(fn != nullptr, "cuMemGetAddressRange driver entry point not available");
    g_get_range = reinterpret_cast<PFN_cuMemGetAddressRange_v3020>(fn);
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  const CUresult r = g_get_range(&base, &size, reinterpret_cast<CUdeviceptr>(ptr));
  if (r != CUDA_SUCCESS) return fail(ETP_ERR_CUDA, "cuMemGetAddressRange failed: " + std::to_string(static_cast<int>(r)));
  ETP_CHECK_CUDA(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle64), reinterpret_cast<void*>(base)));
  *offset = static_cast<int64_t>(reinterpret_cast<CUdeviceptr>(ptr) - base);
  return ETP_OK;
}

ETP_API int etp_ipc_open(const void* handle64, void** base_out) {
  ETP_REQUIRE(handle64 && base_out, "etp_ipc_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  ETP_CHECK_CUDA(cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess));
  return ETP_OK;
}

ETP_API int etp_ipc_close(void* base) {
  if (base) ETP_CHECK_CUDA(cudaIpcCloseMemHandle(base));
  return ETP_OK;
}

ETP_API int etp_peer_signal(const etp_peer_group* g, int32_t kind, int32_t bucket, uint32_t value, void* stream) {
  PeerDev d;
  if (int e = to_dev(g, &d)) return e;
  ETP_REQUIRE((kind == 0 || kind == 1) && bucket >= 0 && bucket < kMaxBuckets, "etp_peer_signal: bad kind / bucket");
  peer_signal_kernel<<<1, 32, 0, S(stream)>>>(d, kind, bucket, value);
  ETP_LAUNCHED();
  return ETP_OK;
}

ETP_API int etp_peer_wait(const etp_peer_group* g, int32_t kind, int32_t bucket_lo, int32_t bucket_hi, uint32_t value,
                          double timeout_s, void* stream) {
  PeerDev d;
  if (int e = to_dev(g, &d)) return e;
  ETP_REQUIRE((kind == 0 || kind == 1) && bucket_lo >= 0 && bucket_lo < bucket_hi && bucket_hi <= kMaxBuckets,
              "etp_peer_wait: bad kind / bucket range");
  const double t = timeout_s > 0 ? timeout_s : 10.0;
  peer_wait_kernel<<<1, 256, 0, S(stream)>>>(d, kind, bucket_lo, bucket_hi, value, static_cast<unsigned long long>(t * 1e9));
  ETP_LAUNCHED();
  return ETP_OK;
}

ETP_API int etp_peer_reduce_adamw(const etp_peer_group* g, int64_t offset, int64_t n, float* exp_avg, float* exp_avg_sq,
                                  float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                                  int32_t write_reduced, int32_t ctas, void* stream) {
  PeerDev d;
  if (int e = to_dev(g, &d)) return e;
  ETP_REQUIRE(offset >= 0 && offset % 4 == 0 && n >= 0 && n % 4 == 0, "etp_peer_reduce_adamw: offset / count must be multiples of 4");
  ETP_REQUIRE(step >= 1, "etp_peer_reduce_adamw: step counts from 1");
  if (n == 0) return ETP_OK;
  ETP_REQUIRE(exp_avg && exp_avg_sq, "etp_peer_reduce_adamw: null optimizer state");
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  int64_t blocks = (n / 4 + 255) / 256;
  const int64_t cap = ctas > 0 ? ctas : 64;
  if (blocks > cap) blocks = cap;
  const float decay = 1.0f - lr * weight_decay;
  const float gscale = 1.0f / static_cast<float>(d.world);
  auto* m4 = reinterpret_cast<float4*>(exp_avg);
  auto* v4 = reinterpret_cast<float4*>(exp_avg_sq);
#define ETP_PEER_LAUNCH(W, U)                                                                                                  \
  peer_reduce_adamw_kernel<W, U><<<static_cast<int>(blocks), 256, 0, S(stream)>>>(d, offset / 4, n / 4, m4, v4, lr, beta1, beta2, \
                                                                                  eps, decay, bc1, sqrtf(bc2), gscale, write_reduced)
  if (d.world <= 2) ETP_PEER_LAUNCH(2, 8);
  else if (d.world <= 4) ETP_PEER_LAUNCH(4, 4);
  else ETP_PEER_LAUNCH(8, 2);
#undef ETP_PEER_LAUNCH
  ETP_LAUNCHED();
  return ETP_OK;
}

ETP_API int etp_peer_error(int32_t* out) {
  ETP_REQUIRE(out != nullptr, "etp_peer_error: null argument");
  int v = 0;
  ETP_CHECK_CUDA(cudaMemcpyFromSymbol(&v, g_peer_error, sizeof(int)));
  *out = v;
  return ETP_OK;
}

}  // extern "C"
