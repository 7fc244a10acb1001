// Segmented weighted row gather (CSR): the one new kernel of the pre-training twin.
//   out[s, :] = sum_{k in [seg_ptr[s], seg_ptr[s+1])} weight[k] * src[index[k], :]
// It is the forward of GlobalMapEncoder._aggregate_gmap_features
// (pretrain_src/pretrain_src/model/vilmodel.py:585-619: a visited viewpoint is the mean of its valid view tokens, an
// unvisited one the mean of the candidate-view tokens that pointed at it; the host turns the reference's string-keyed
// dictionaries into the CSR, etpnav_b200/pretrain.py), its backward (same kernel over the transposed structure, so
// no atomics and a deterministic sum), and the masked-token gather of _compute_masked_hidden
// (pretrain_cmt.py:160-164) with unit weights.
// HBM-bound: one CTA per output row, one float4 column per thread (768 floats = 192 threads), entries walked in
// order; algorithmic bytes = (nnz + segments) * width * 4.
#include "../../include/etpnav_b200.h"
#include "common.cuh"
#include "ops.h"

namespace etp {
namespace {

__global__ void segment_gather_kernel(const float* __restrict__ src, const int32_t* __restrict__ seg_ptr,
                                      const int32_t* __restrict__ index, const float* __restrict__ weight, int width4,
                                      float* __restrict__ out) {
  griddep_launch();  // PDL (common.cuh): let the next kernel get scheduled ...
  griddep_wait();    // ... and wait for the previous one before touching memory
  const int s = blockIdx.x;
  const int beg = seg_ptr[s], end = seg_ptr[s + 1];
  for (int c = threadIdx.x; c < width4; c += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = beg; k < end; ++k) {
      const float w = weight[k];
      const float4 v = reinterpret_cast<const float4*>(src + static_cast<size_t>(index[k]) * width4 * 4)[c];
      acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
    }
    reinterpret_cast<float4*>(out + static_cast<size_t>(s) * width4 * 4)[c] = acc;
  }
}

// same gather with every source row at its own address (a table of row pointers): the map keeps one tensor per node
// (GraphMap.node_embeds / ghost_embeds, graph_utils.py:137-144), so the inference path reads them where they are
// instead of stacking them first
__global__ void segment_gather_rows_kernel(const float* const* __restrict__ rows, const int32_t* __restrict__ seg_ptr,
                                           const int32_t* __restrict__ index, const float* __restrict__ weight, int width4,
                                           float* __restrict__ out) {
  griddep_launch();  // PDL (common.cuh): let the next kernel get scheduled ...
  griddep_wait();    // ... and wait for the previous one before touching memory
  const int s = blockIdx.x;
  const int beg = seg_ptr[s], end = seg_ptr[s + 1];
  for (int c = threadIdx.x; c < width4; c += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = beg; k < end; ++k) {
      const float w = weight[k];
      const float4 v = reinterpret_cast<const float4*>(rows[index[k]])[c];
      acc.x += w * v.x; acc.y += w * v.y; acc.z += w * v.z; acc.w += w * v.w;
    }
    reinterpret_cast<float4*>(out + static_cast<size_t>(s) * width4 * 4)[c] = acc;
  }
}

}  // namespace

int segment_gather_rows(const float* const* rows, const int32_t* seg_ptr, const int32_t* index, const float* weight,
                        int num_segments, int width, float* out, cudaStream_t stream) {
  if (num_segments <= 0) return ETP_OK;
  ETP_REQUIRE(rows && seg_ptr && index && weight && out, "segment_gather_rows: null argument");
  ETP_REQUIRE(width > 0 && width % 4 == 0, "segment_gather_rows: width must be a multiple of 4");
  ETP_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, "segment_gather_rows: output must be 16-byte aligned");
  const int width4 = width / 4;
  int threads = (width4 + 31) / 32 * 32;
  if (threads > 256) threads = 256;
  ETP_CHECK_CUDA(launch_pdl(segment_gather_rows_kernel, dim3(num_segments), dim3(threads), 0, stream, rows, seg_ptr, index,
                            weight, width4, out));
  ETP_LAUNCHED();
  return ETP_OK;
}

int segment_gather(const float* src, const int32_t* seg_ptr, const int32_t* index, const float* weight, int num_segments,
                   int width, float* out, cudaStream_t stream) {
  if (num_segments <= 0) return ETP_OK;
  ETP_REQUIRE(src && seg_ptr && index && weight && out, "segment_gather: null argument");
  ETP_REQUIRE(width > 0 && width % 4 == 0, "segment_gather: width must be a multiple of 4");
  ETP_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
              "segment_gather: rows must be 16-byte aligned");
  const int width4 = width / 4;
  int threads = (width4 + 31) / 32 * 32;
  if (threads > 256) threads = 256;
  ETP_CHECK_CUDA(launch_pdl(segment_gather_kernel, dim3(num_segments), dim3(threads), 0, stream, src, seg_ptr, index,
                            weight, width4, out));
  ETP_LAUNCHED();
  return ETP_OK;
}

}  // namespace etp

extern "C" __attribute__((visibility("default"))) int etp_segment_gather(const float* src, const int32_t* seg_ptr,
                                                                          const int32_t* index, const float* weight,
                                                                          int32_t num_segments, int32_t width, float* out,
                                                                          void* stream) {
  return etp::segment_gather(src, seg_ptr, index, weight, num_segments, width, out, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" __attribute__((visibility("default"))) int etp_segment_gather_rows(const float* const* rows, const int32_t* seg_ptr,
                                                                               const int32_t* index, const float* weight,
                                                                               int32_t num_segments, int32_t width,
                                                                               float* out, void* stream) {
  return etp::segment_gather_rows(rows, seg_ptr, index, weight, num_segments, width, out,
                                  reinterpret_cast<cudaStream_t>(stream));
}
