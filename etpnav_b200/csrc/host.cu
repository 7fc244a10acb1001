#include "host.h"

#include <cudaTypedefs.h>

#include <cstring>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace etp {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
const char* last_error_cstr() { return g_last_error.c_str(); }

// ------------------------------------------------------------------------------------------------
// cuTensorMapEncodeTiled through cudaGetDriverEntryPoint: no link-time dependency on libcuda.so, so
// the library can be dlopen'ed (and its exports checked) on a machine without a driver.
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::mutex g_mu;

static int load_encode() {
  if (g_encode) return ETP_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr)
    return fail(ETP_ERR_CUDA, "cuTensorMapEncodeTiled driver entry point not available (no CUDA driver?)");
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return ETP_OK;
}

struct TmapKey {
  const void* ptr;
  uint64_t a, b, c, d, e;
  uint32_t b0, b1, rank;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && a == o.a && b == o.b && c == o.c && d == o.d && e == o.e && b0 == o.b0 && b1 == o.b1 &&
           rank == o.rank;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.a); mix(k.b); mix(k.c); mix(k.d); mix(k.e); mix(k.b0); mix(k.b1); mix(k.rank);
    return h;
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_cache;

int get_tmap_2d(const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows, uint32_t box_cols,
                CUtensorMap* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  int rc = load_encode();
  if (rc) return rc;
  TmapKey key{ptr, rows, cols, ld, 0, 0, box_rows, box_cols, 2};
  auto it = g_cache.find(key);
  if (it != g_cache.end()) {
    *out = it->second;
    return ETP_OK;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 2) & 15))
    return fail(ETP_ERR_INVALID, "TMA operand must be 16-byte aligned with a 16-byte-multiple row pitch");
  if (box_cols * 2 != 128 || box_rows > 256) return fail(ETP_ERR_INVALID, "TMA box must be 128 B wide, <= 256 rows");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ETP_ERR_CUDA, "cuTensorMapEncodeTiled(2d) failed: " + std::to_string((int)r));
  if (g_cache.size() > 65536) g_cache.clear();
  g_cache.emplace(key, m);
  *out = m;
  return ETP_OK;
}

int get_tmap_3d(const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1, uint64_t ld2, uint32_t box0,
                uint32_t box1, CUtensorMap* out) {
  std::lock_guard<std::mutex> lk(g_mu);
  int rc = load_encode();
  if (rc) return rc;
  TmapKey key{ptr, d0, d1, d2, ld1, ld2, box0, box1, 3};
  auto it = g_cache.find(key);
  if (it != g_cache.end()) {
    *out = it->second;
    return ETP_OK;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld1 * 2) & 15) || ((ld2 * 2) & 15))
    return fail(ETP_ERR_INVALID, "TMA operand must be 16-byte aligned with 16-byte-multiple pitches");
  if (box0 * 2 != 128 || box1 > 256) return fail(ETP_ERR_INVALID, "TMA box must be 128 B wide, <= 256 rows");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {ld1 * 2, ld2 * 2};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap m;
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ETP_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed: " + std::to_string((int)r));
  if (g_cache.size() > 65536) g_cache.clear();
  g_cache.emplace(key, m);
  *out = m;
  return ETP_OK;
}

std::atomic<long long> g_launches{0};

// ---- per-launch profiling (off by default) ----
bool g_prof_on = false;
struct ProfRec { cudaEvent_t a, b; const void* func; std::string tag; double flops; };
static std::vector<ProfRec> g_prof;
static cudaEvent_t g_prof_cur = nullptr;
static std::string g_prof_tag;
static double g_prof_flops = 0.0;

void prof_tag(const char* tag, double flops) {
  if (!g_prof_on) return;
  g_prof_tag = tag ? tag : "";
  g_prof_flops = flops;
}
void prof_begin(cudaStream_t s) {
  cudaEventCreate(&g_prof_cur);
  cudaEventRecord(g_prof_cur, s);
}
void prof_end(cudaStream_t s, const void* func) {
  if (!g_prof_cur) return;
  ProfRec r;
  r.a = g_prof_cur;
  cudaEventCreate(&r.b);
  cudaEventRecord(r.b, s);
  r.func = func;
  r.tag.swap(g_prof_tag);
  r.flops = g_prof_flops;
  g_prof_flops = 0.0;
  g_prof.push_back(std::move(r));
  g_prof_cur = nullptr;
}
void prof_enable(bool on) { g_prof_on = on; }

static std::string func_name(const void* f) {
  const char* n = nullptr;
  if (cudaFuncGetName(&n, f) != cudaSuccess || !n) return "?";
  return n;
}

// Totals over the launches that carried FLOPs (the tcgen05 GEMM); optionally a text report, one line per
// (kernel, tag): "<count>\t<total_ms>\t<flops>\t<kernel>\t<tag>".  Clears the records.
int prof_collect(double* ms, double* flops, long long* count, char* report, size_t cap) {
  double t = 0, f = 0;
  long long n = 0;
  struct Agg { long long n = 0; double ms = 0, fl = 0; };
  std::unordered_map<std::string, Agg> agg;
  std::vector<std::string> order;
  for (auto& r : g_prof) {
    if (cudaEventSynchronize(r.b) != cudaSuccess) return fail(ETP_ERR_CUDA, "prof_collect: event sync failed");
    float e = 0;
    cudaEventElapsedTime(&e, r.a, r.b);
    if (r.flops > 0) { t += e; f += r.flops; ++n; }
    if (report) {
      std::string key = func_name(r.func) + "\t" + r.tag;
      auto it = agg.find(key);
      if (it == agg.end()) { order.push_back(key); it = agg.emplace(key, Agg()).first; }
      it->second.n++; it->second.ms += e; it->second.fl += r.flops;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  if (ms) *ms = t;
  if (flops) *flops = f;
  if (count) *count = n;
  if (report && cap > 0) {
    std::string out;
    for (auto& k : order) {
      const Agg& a = agg[k];
      out += std::to_string(a.n) + "\t" + std::to_string(a.ms) + "\t" + std::to_string(a.fl) + "\t" + k + "\n";
    }
    if (out.size() >= cap) out.resize(cap - 1);
    memcpy(report, out.c_str(), out.size() + 1);
  }
  g_prof.clear();
  return ETP_OK;
}

// SMs the persistent grids (GEMM pairs, attention, row kernels) size themselves for.  A data-parallel host can hold a few
// SMs back for the collective's CTAs (etp_set_sm_reserve): a persistent all-SM cluster grid otherwise waits, tile after
// tile, for the SM pairs an NCCL kernel occupies, and the all-reduce waits for SMs the GEMMs never release.
static std::atomic<int> g_sm_reserve{0};
void set_sm_reserve(int n) { g_sm_reserve.store(n < 0 ? 0 : n); }
int get_sm_reserve() { return g_sm_reserve.load(); }
int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  int avail = n - g_sm_reserve.load(std::memory_order_relaxed);
  avail &= ~1;              // whole SM pairs (cta_group::2 clusters)
  return avail < 2 ? 2 : avail;
}

}  // namespace etp
