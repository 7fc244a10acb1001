// Geometry of the per-step topological-map tensors, as functions usable on the device (the
// etp_gmap_pack kernel, gmap_pack.cu) and on the host (tests/host_harness/gmap_pack_host.cu runs the SAME arithmetic on
// the CPU against the oracle; test infrastructure only).  Restates, in double precision with IEEE-exact operation order
// (no FMA contraction, so distances round to the same float32 as numpy's):
//   calc_position_distance        vlnce_baselines/models/graph_utils.py:13-19
//   calculate_vp_rel_pos_fts      graph_utils.py:21-45   (to_clock = True, base_elevation = 0 as get_pos_fts calls it)
//   GraphMap.front_to_ghost_dist  graph_utils.py:258-270 (nearest front, first minimum wins)
//   GraphMap.get_pos_fts          graph_utils.py:278-322
//   pair distances of _nav_gmap_variable   vlnce_baselines/ss_trainer_ETP.py:371-387
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define ETP_HD __host__ __device__ inline
#else
#define ETP_HD inline
#endif

namespace etp {

constexpr double kMaxDist = 30.0;  // graph_utils.py:9
constexpr double kMaxStep = 10.0;  // graph_utils.py:10
constexpr double kPi = 3.141592653589793;

// exact (unfused) a*a + b*b (+ c*c) in the order numpy evaluates dx**2 + dy**2 + dz**2
ETP_HD double sq_sum3(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(__dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b)), __dmul_rn(c, c));
#else
  volatile double aa = a * a, bb = b * b, cc = c * c;
  volatile double s = aa + bb;
  return s + cc;
#endif
}
ETP_HD double sq_sum2(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(__dmul_rn(a, a), __dmul_rn(b, b));
#else
  volatile double aa = a * a, bb = b * b;
  return aa + bb;
#endif
}

ETP_HD double position_distance(const double* a, const double* b) {
  return sqrt(sq_sum3(b[0] - a[0], b[1] - a[1], b[2] - a[2]));
}

// One view of the environment's map, flattened by the host packer (etpnav_b200/packing.py).  Nodes are indexed in the
// order of GraphMap.node_pos, ghosts in the order of GraphMap.ghost_pos (ss_trainer_ETP.py:351-352).
struct GmapEnvView {
  int n_nodes, n_ghosts, cur_node;
  const double* cur_pos;       // [3]
  double base_heading;         // heading_from_quaternion(cur_ori), graph_utils.py:53-58
  const double* node_pos;      // [n_nodes,3]
  const double* ghost_pos;     // [n_ghosts,3]  (ghost_aug_pos)
  const double* dist;          // [n_nodes,n_nodes] shortest_dist between nodes
  const int32_t* node_step;    // [n_nodes] node_stepId
  const int32_t* front_ptr;    // [n_ghosts+1] CSR of ghost_fronts (node indices)
  const int32_t* front_idx;
  const int32_t* path_len;     // [n_nodes,n_nodes] len(shortest_path[a][b])
};

// front_to_ghost_dist: nearest front of ghost g
ETP_HD void ghost_front(const GmapEnvView& e, int g, double* dis, int* front) {
  double best = 10000.0;
  int bf = -1;
  for (int k = e.front_ptr[g]; k < e.front_ptr[g + 1]; ++k) {
    const int f = e.front_idx[k];
    const double d = position_distance(e.node_pos + 3 * f, e.ghost_pos + 3 * g);
    if (d < best) { best = d; bf = f; }
  }
  *dis = best;
  *front = bf;
}

// row r of gmap_pos_fts (7 floats): r = 0 [stop], 1..n nodes, n+1..n+g ghosts (front data precomputed)
ETP_HD void pos_fts_row(const GmapEnvView& e, int r, const double* front_dis, const int* front, float* out) {
  if (r == 0) {  // vp is None: angles 0 -> (sin 0, cos 0, sin 0, cos 0), dists 0 (graph_utils.py:284-286)
    out[0] = 0.f; out[1] = 1.f; out[2] = 0.f; out[3] = 1.f; out[4] = 0.f; out[5] = 0.f; out[6] = 0.f;
    return;
  }
  const bool ghost = r > e.n_nodes;
  const int k = ghost ? r - 1 - e.n_nodes : r - 1;
  const double* b = ghost ? e.ghost_pos + 3 * k : e.node_pos + 3 * k;
  const double* a = e.cur_pos;
  const double dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
  double xz = sqrt(sq_sum2(dx, dz));
  if (!(xz > 1e-8)) xz = 1e-8;
  double xyz = sqrt(sq_sum3(dx, dy, dz));
  if (!(xyz > 1e-8)) xyz = 1e-8;
  double heading = asin(-dx / xz);
  if (b[2] > a[2]) heading = kPi - heading;
  heading -= e.base_heading;
  heading = 2.0 * kPi - heading;  // to_clock
  const double elevation = asin(dz / xyz);
  double sdist, steps;
  if (ghost) {
    sdist = e.dist[static_cast<size_t>(e.cur_node) * e.n_nodes + front[k]] + front_dis[k];
    steps = e.path_len[static_cast<size_t>(e.cur_node) * e.n_nodes + front[k]] + 1;
  } else {
    sdist = e.dist[static_cast<size_t>(e.cur_node) * e.n_nodes + k];
    steps = e.path_len[static_cast<size_t>(e.cur_node) * e.n_nodes + k];
  }
  // rel_angles are rounded to float32 BEFORE sin / cos (np.array(...).astype(np.float32), :318; get_angle_fts :47-53)
  const float hf = static_cast<float>(heading), ef = static_cast<float>(elevation);
  out[0] = sinf(hf); out[1] = cosf(hf); out[2] = sinf(ef); out[3] = cosf(ef);
  out[4] = static_cast<float>(xyz / kMaxDist);
  out[5] = static_cast<float>(sdist / kMaxDist);
  out[6] = static_cast<float>(steps / kMaxStep);
}

// gmap_pair_dists[r, c] for 0 <= r, c < len (ss_trainer_ETP.py:371-387): evaluated with j = min, k = max as the
// reference's upper-triangle loop does, so both halves are the same float
ETP_HD float pair_dist(const GmapEnvView& e, int r, int c, const double* front_dis, const int* front) {
  if (r == c || r == 0 || c == 0) return 0.f;
  const int j = r < c ? r : c, k = r < c ? c : r;
  const bool gj = j > e.n_nodes, gk = k > e.n_nodes;
  const int a = gj ? j - 1 - e.n_nodes : j - 1, b = gk ? k - 1 - e.n_nodes : k - 1;
  double d;
  if (!gj && !gk) {
    d = e.dist[static_cast<size_t>(a) * e.n_nodes + b];
  } else if (!gj && gk) {
    d = e.dist[static_cast<size_t>(a) * e.n_nodes + front[b]] + front_dis[b];
  } else {  // both ghosts (nodes always precede ghosts, so gj implies gk)
    d = front_dis[a] + e.dist[static_cast<size_t>(front[a]) * e.n_nodes + front[b]];
    d = d + front_dis[b];
  }
  return static_cast<float>(d / kMaxDist);
}

}  // namespace etp
