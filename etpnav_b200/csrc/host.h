// Host-side plumbing shared by the launchers: error reporting for the C ABI, the TMA tensor-map
// encoder (driver entry point resolved through cudart so the library loads without libcuda.so).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <utility>
#include <string>

namespace etp {

// error codes (same values as include/etpnav_b200.h)
#ifndef ETP_OK
#define ETP_OK 0
#define ETP_ERR_INVALID (-1)   // bad argument / unsupported shape
#define ETP_ERR_CUDA (-2)      // CUDA runtime / driver error
#define ETP_ERR_NO_DEVICE (-3) // no sm_100 device
#endif

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define ETP_CHECK_CUDA(expr)                                                                       \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::etp::fail(ETP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// every kernel launch site ends with this: counts the launch (bench.py reports gpu_launches) and checks it
extern std::atomic<long long> g_launches;
#define ETP_LAUNCHED()                                            \
  do {                                                            \
    ::etp::g_launches.fetch_add(1, std::memory_order_relaxed);    \
    ETP_CHECK_CUDA(cudaGetLastError());                           \
  } while (0)

// launch with programmatic stream serialization (see griddep_* in common.cuh)
// optional per-launch CUDA-event timing of every kernel (bench.py roofline leg, profiles/): when enabled,
// launch_pdl brackets each launch with a pair of events; prof_tag() labels the next launch (GEMM shape, FLOPs).
extern bool g_prof_on;
void prof_tag(const char* tag, double flops);
void prof_begin(cudaStream_t s);
void prof_end(cudaStream_t s, const void* func);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (g_prof_on) prof_begin(s);
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
  if (g_prof_on) prof_end(s, reinterpret_cast<const void*>(kern));
  return e;
}

#define ETP_REQUIRE(cond, msg)                                                     \
  do {                                                                             \
    if (!(cond)) return ::etp::fail(ETP_ERR_INVALID, std::string(msg));     \
  } while (0)

// 2-D bf16 tensor map: global [rows, cols] with row pitch ld (elements), box [box_rows, box_cols],
// 128-byte swizzle (box_cols * 2 bytes must be 128).  Cached by (ptr, shape, box).
int get_tmap_2d(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols,
                CUtensorMap* out);
// 3-D bf16 tensor map: global [d2, d1, d0] (d0 contiguous) with pitches ld1 (elements between d1 rows),
// ld2 (elements between d2 slabs); box [1, box1, box0]; 128-byte swizzle.
int get_tmap_3d(const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1, uint64_t ld2, uint32_t box0,
                uint32_t box1, CUtensorMap* out);

int num_sms();               // SMs available to the library's persistent grids (device SMs minus the reserve)
void set_sm_reserve(int n);
int get_sm_reserve();  // SMs left to other kernels (collectives) while data-parallel training

// host side of the dropout scheme (common.cuh: Drop): key of a site from the call's seed (splitmix64), threshold and
// scale from the probability.  p <= 0 gives thr = 0 (off).
struct DropHost { uint32_t key, thr; float scale; };
inline DropHost make_drop_host(uint64_t seed, float p, uint32_t site) {
  DropHost d{0u, 0u, 1.0f};
  if (!(p > 0.0f)) return d;
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (static_cast<uint64_t>(site) + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  d.key = static_cast<uint32_t>(z);
  double t = static_cast<double>(p) * 65536.0 + 0.5;
  if (t > 65535.0) t = 65535.0;
  d.thr = static_cast<uint32_t>(t);
  if (d.thr == 0) d.thr = 1;
  d.scale = 1.0f / (1.0f - p);
  return d;
}

}  // namespace etp
