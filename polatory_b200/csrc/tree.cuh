// Device octree (1-D / quadtree / octree) over Morton-sorted points.
//
// Replaces, on the device, what the reference gets from ScalFMM:
//   scalfmm::utils::sort_container(box, level, particles)   src/fmm/fmm_evaluator.hpp:239-246
//   scalfmm::component::group_tree_view(height, order, box, 10, 10, particles, sorted)
//                                                            src/fmm/fmm_evaluator.hpp:260-268
// Layout: a *uniform* tree of `height` levels (0 .. height-1, leaves at height-1, as
// src/fmm/utility.hpp:12-16 fixes the height from the point count).  Because the height
// rule keeps the number of leaf cells proportional to the number of points, every level
// keeps a dense key -> compact-cell map (int32, 2^(dim*level) entries); neighbour and
// interaction-list lookups are O(1) loads instead of searches.
#pragma once

#include <cstdint>
#include <memory>
#include <vector>

#include "common.cuh"

namespace plt {

// ---- Morton helpers (host + device) ---------------------------------------------------
template <int DIM>
__host__ __device__ __forceinline__ uint32_t morton_encode(const int (&c)[DIM]) {
  if constexpr (DIM == 1) {
    return static_cast<uint32_t>(c[0]);
  } else if constexpr (DIM == 2) {
    auto spread = [](uint32_t v) {
      v &= 0xffffu;
      v = (v | (v << 8)) & 0x00ff00ffu;
      v = (v | (v << 4)) & 0x0f0f0f0fu;
      v = (v | (v << 2)) & 0x33333333u;
      v = (v | (v << 1)) & 0x55555555u;
      return v;
    };
    return (spread(c[0]) << 1) | spread(c[1]);
  } else {
    auto spread = [](uint32_t v) {
      v &= 0x3ffu;
      v = (v | (v << 16)) & 0x030000ffu;
      v = (v | (v << 8)) & 0x0300f00fu;
      v = (v | (v << 4)) & 0x030c30c3u;
      v = (v | (v << 2)) & 0x09249249u;
      return v;
    };
    return (spread(c[0]) << 2) | (spread(c[1]) << 1) | spread(c[2]);
  }
}

template <int DIM>
__host__ __device__ __forceinline__ void morton_decode(uint32_t key, int (&c)[DIM]) {
  if constexpr (DIM == 1) {
    c[0] = static_cast<int>(key);
  } else if constexpr (DIM == 2) {
    auto compact = [](uint32_t v) {
      v &= 0x55555555u;
      v = (v | (v >> 1)) & 0x33333333u;
      v = (v | (v >> 2)) & 0x0f0f0f0fu;
      v = (v | (v >> 4)) & 0x00ff00ffu;
      v = (v | (v >> 8)) & 0x0000ffffu;
      return v;
    };
    c[0] = static_cast<int>(compact(key >> 1));
    c[1] = static_cast<int>(compact(key));
  } else {
    auto compact = [](uint32_t v) {
      v &= 0x09249249u;
      v = (v | (v >> 2)) & 0x030c30c3u;
      v = (v | (v >> 4)) & 0x0300f00fu;
      v = (v | (v >> 8)) & 0x030000ffu;
      v = (v | (v >> 16)) & 0x000003ffu;
      return v;
    };
    c[0] = static_cast<int>(compact(key >> 2));
    c[1] = static_cast<int>(compact(key >> 1));
    c[2] = static_cast<int>(compact(key));
  }
}

// Root box: src/fmm/utility.hpp:18-33 (cube, width 1.01 * max extent of the transformed
// bbox, centred on it).
struct Box {
  double width = 1.0;
  double center[kMaxDim] = {0, 0, 0};
};

// Device view of one tree, passed by value to kernels.
struct TreeView {
  int dim;
  int height;          // levels 0 .. height-1
  int64_t n;           // points
  const double* pos;   // SoA [dim][n], Morton-sorted, anisotropy-transformed
  const int* perm;     // sorted index -> caller index
  const int* leaf_start;  // [n_leaf + 1] point ranges of the compact leaves
  // Per level: offset of the level inside the concatenated arrays, and cell counts.
  const int* dense;      // concatenated dense maps; level l starts at dense_off[l]
  const uint32_t* keys;  // concatenated compact Morton keys; level l starts at cell_off[l]
  int64_t dense_off[24];
  int cell_off[24];
  int n_cells[24];
  // Optional per-cell work flags of a partitioned (multi-GPU) upward pass, indexed by the global compact cell
  // id: bit 0 = this rank computes the cell's multipole expansion M (P2M / M2M), bit 1 = its spectrum Mhat.
  // nullptr = every cell (single GPU).
  const unsigned char* flags = nullptr;
};
constexpr unsigned char kCellFlagM = 1, kCellFlagMhat = 2;

class Tree {
 public:
  // pos_caller: SoA [dim][n] transformed positions in caller order (device).
  void build(int dim, int height, const Box& box, const double* pos_caller, int64_t n,
             cudaStream_t stream, LaunchCounter& ctr);
  bool built() const { return height_ > 0; }
  void reset() { height_ = 0; }
  int height() const { return height_; }
  int64_t n() const { return n_; }
  int n_cells(int level) const { return n_cells_[level]; }
  int total_cells() const { return total_cells_; }
  TreeView view() const;
  const double* pos() const { return pos_.get(); }
  const int* perm() const { return perm_.get(); }

 private:
  int dim_ = 0, height_ = 0;
  int64_t n_ = 0;
  DevBuf<double> pos_;
  DevBuf<int> perm_;
  DevBuf<uint32_t> pkey_;
  DevBuf<int> dense_;
  DevBuf<uint32_t> keys_;
  DevBuf<int> leaf_start_;
  DevBuf<unsigned char> tmp_;
  DevBuf<uint32_t> key_tmp_;
  DevBuf<int> idx_tmp_;
  DevBuf<int> occ_, scan_;
  DevBuf<int64_t> d_off_;
  DevBuf<int> summary_;
  struct HostFree {
    void operator()(int* p) const { cudaFreeHost(p); }
  };
  std::unique_ptr<int, HostFree> h_summary_;  // pinned landing buffer of the level summary
  std::vector<int64_t> dense_off_;
  std::vector<int> cell_off_, n_cells_;
  int total_cells_ = 0;
};

// src/fmm/utility.hpp:12-16.
int fmm_tree_height(int dim, int64_t n_points);

}  // namespace plt
