// etp_gmap_pack: the numeric half of the trainer's per-step map packing (_nav_gmap_variable, ss_trainer_ETP.py:344-417,
// with GraphMap.get_pos_fts, graph_utils.py:278-322) as ONE launch over the batch of environments, writing the padded
// device tensors the planner consumes directly: gmap_step_ids, gmap_visited_masks, gmap_masks, gmap_pos_fts,
// gmap_pair_dists.  The reference builds them in Python (O(B.N^2) loops over string-keyed dictionaries, one .cuda() per
// tensor); here the host flattens the map state into one pinned blob (etpnav_b200/packing.py), one H2D copy brings it
// over, and the kernel does the geometry.  HBM-bound: the output is B.N.(N + 7) floats + 10 bytes per node.
// One CTA per environment: ghost fronts -> shared memory, then rows, then the pair matrix.
#include "../../include/etpnav_b200.h"
#include "common.cuh"
#include "gmap_pack.cuh"
#include "host.h"

namespace etp {
namespace {

constexpr int kMaxGhosts = 1024;

// meta[env]: {n_nodes, n_ghosts, cur_node, off_f64, off_i32, nnz_fronts, 0, 0}; offsets in elements into the two blobs.
// f64 blob per env: cur_pos[3], base_heading, node_pos[3n], ghost_pos[3g], dist[n*n]
// i32 blob per env: node_step[n], front_ptr[g+1], front_idx[nnz], path_len[n*n]
__global__ void gmap_pack_kernel(const int32_t* __restrict__ meta, const double* __restrict__ f64, const int32_t* __restrict__ i32,
                                 int n_max, int64_t* __restrict__ step_ids, uint8_t* __restrict__ visited,
                                 uint8_t* __restrict__ masks, float* __restrict__ pos_fts, float* __restrict__ pair_dists) {
  griddep_launch();  // PDL (common.cuh): let the next kernel get scheduled ...
  griddep_wait();    // ... and wait for the previous one before touching memory
  __shared__ double s_front_dis[kMaxGhosts];
  __shared__ int s_front[kMaxGhosts];
  const int b = blockIdx.x;
  const int32_t* m = meta + 8 * b;
  GmapEnvView e;
  e.n_nodes = m[0]; e.n_ghosts = m[1]; e.cur_node = m[2];
  const double* d = f64 + m[3];
  const int32_t* q = i32 + m[4];
  e.cur_pos = d; e.base_heading = d[3];
  e.node_pos = d + 4;
  e.ghost_pos = e.node_pos + 3 * e.n_nodes;
  e.dist = e.ghost_pos + 3 * e.n_ghosts;
  e.node_step = q;
  e.front_ptr = q + e.n_nodes;
  e.front_idx = e.front_ptr + e.n_ghosts + 1;
  e.path_len = e.front_idx + m[5];
  const int len = 1 + e.n_nodes + e.n_ghosts;
  for (int g = threadIdx.x; g < e.n_ghosts; g += blockDim.x) ghost_front(e, g, &s_front_dis[g], &s_front[g]);
  __syncthreads();
  for (int r = threadIdx.x; r < n_max; r += blockDim.x) {
    const size_t o = static_cast<size_t>(b) * n_max + r;
    float row[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int64_t sid = 0;
    uint8_t vis = 0;
    if (r < len) {
      pos_fts_row(e, r, s_front_dis, s_front, row);
      if (r >= 1 && r <= e.n_nodes) { sid = e.node_step[r - 1]; vis = 1; }
    }
    step_ids[o] = sid;
    visited[o] = vis;
    masks[o] = r < len ? 1 : 0;
#pragma unroll
    for (int k = 0; k < 7; ++k) pos_fts[o * 7 + k] = row[k];
  }
  float* pd = pair_dists + static_cast<size_t>(b) * n_max * n_max;
  for (int i = threadIdx.x; i < n_max * n_max; i += blockDim.x) {
    const int r = i / n_max, c = i - r * n_max;
    pd[i] = (r < len && c < len) ? pair_dist(e, r, c, s_front_dis, s_front) : 0.f;
  }
}

}  // namespace

int gmap_pack(const int32_t* meta, const double* f64, const int32_t* i32, int B, int n_max, int max_ghosts, int64_t* step_ids,
              uint8_t* visited, uint8_t* masks, float* pos_fts, float* pair_dists, cudaStream_t stream) {
  if (B <= 0) return ETP_OK;
  ETP_REQUIRE(meta && f64 && i32 && step_ids && visited && masks && pos_fts && pair_dists, "gmap_pack: null argument");
  ETP_REQUIRE(n_max >= 1 && n_max <= 2048, "gmap_pack: padded map size out of range");
  ETP_REQUIRE(max_ghosts >= 0 && max_ghosts <= kMaxGhosts, "gmap_pack: at most 1024 ghost nodes per environment");
  ETP_CHECK_CUDA(launch_pdl(gmap_pack_kernel, dim3(B), dim3(256), 0, stream, meta, f64, i32, n_max, step_ids, visited, masks,
                            pos_fts, pair_dists));
  ETP_LAUNCHED();
  return ETP_OK;
}

}  // namespace etp

extern "C" __attribute__((visibility("default"))) int etp_gmap_pack(const int32_t* meta, const double* f64_blob,
                                                                     const int32_t* i32_blob, int32_t B, int32_t n_max,
                                                                     int32_t max_ghosts, int64_t* gmap_step_ids,
                                                                     uint8_t* gmap_visited_masks, uint8_t* gmap_masks,
                                                                     float* gmap_pos_fts, float* gmap_pair_dists, void* stream) {
  return etp::gmap_pack(meta, f64_blob, i32_blob, B, n_max, max_ghosts, gmap_step_ids, gmap_visited_masks, gmap_masks,
                        gmap_pos_fts, gmap_pair_dists, reinterpret_cast<cudaStream_t>(stream));
}
