// TEST INFRASTRUCTURE: runs the geometry of etpnav_b200/csrc/gmap_pack.cuh (the functions the etp_gmap_pack kernel
// calls) on the CPU, over the same blob layout, so the CPU test-suite can check the device arithmetic and the host
// packer's flattening against the reference fixtures without a GPU.  Compiled by tests/test_packing_host_cpu.py with g++
// into tests/_build/ (never part of libetpnav_b200.so: the product has no CPU path).
#include <vector>

#include "../../etpnav_b200/csrc/gmap_pack.cuh"

extern "C" int gmap_pack_host(const int32_t* meta, const double* f64, const int32_t* i32, int B, int n_max, int64_t* step_ids,
                              uint8_t* visited, uint8_t* masks, float* pos_fts, float* pair_dists) {
  using namespace etp;
  for (int b = 0; b < B; ++b) {
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
    std::vector<double> fd(e.n_ghosts + 1);
    std::vector<int> fr(e.n_ghosts + 1);
    for (int g = 0; g < e.n_ghosts; ++g) ghost_front(e, g, &fd[g], &fr[g]);
    for (int r = 0; r < n_max; ++r) {
      const size_t o = static_cast<size_t>(b) * n_max + r;
      float row[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      int64_t sid = 0;
      uint8_t vis = 0;
      if (r < len) {
        pos_fts_row(e, r, fd.data(), fr.data(), row);
        if (r >= 1 && r <= e.n_nodes) { sid = e.node_step[r - 1]; vis = 1; }
      }
      step_ids[o] = sid;
      visited[o] = vis;
      masks[o] = r < len ? 1 : 0;
      for (int k = 0; k < 7; ++k) pos_fts[o * 7 + k] = row[k];
    }
    float* pd = pair_dists + static_cast<size_t>(b) * n_max * n_max;
    for (int i = 0; i < n_max * n_max; ++i) {
      const int r = i / n_max, c = i - r * n_max;
      pd[i] = (r < len && c < len) ? pair_dist(e, r, c, fd.data(), fr.data()) : 0.f;
    }
  }
  return 0;
}
