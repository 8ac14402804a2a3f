// Device tree build: Morton keys -> radix sort -> dense/compact cell maps for all levels.
#include <cub/cub.cuh>

#include <cmath>

#include "tree.cuh"

namespace plt {

int fmm_tree_height(int dim, int64_t n_points) {
  // src/fmm/utility.hpp:12-16: max(2, round(ln n / ln 2^dim)).
  double h = std::round(std::log(static_cast<double>(n_points)) / std::log(std::pow(2.0, dim)));
  return std::max(2, static_cast<int>(h));
}

namespace {

template <int DIM>
__global__ void k_point_keys(const double* __restrict__ pos, int64_t n, Box box, int level,
                             uint32_t* __restrict__ keys, int* __restrict__ idx, int* __restrict__ outside) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int nside = 1 << level;
  const double inv_w = static_cast<double>(nside) / box.width;
  int c[DIM];
#pragma unroll
  for (int a = 0; a < DIM; ++a) {
    double corner = box.center[a] - 0.5 * box.width;
    const double x = (pos[a * n + i] - corner) * inv_w;
    // The root box is 1.01 x the bbox the evaluator was made with (src/fmm/utility.hpp:18-33), so points of
    // that bbox are strictly inside.  A point outside it (NaN included) would be clamped into a boundary
    // cell whose expansion does not cover it -- silent, unbounded error; reported instead.
    if (!(x >= 0.0 && x <= static_cast<double>(nside))) *outside = 1;
    int ci = static_cast<int>(floor(x));
    c[a] = min(max(ci, 0), nside - 1);
  }
  keys[i] = morton_encode<DIM>(c);
  idx[i] = static_cast<int>(i);
}

__global__ void k_gather_pos(const double* __restrict__ src, const int* __restrict__ perm, int64_t n,
                             int dim, double* __restrict__ dst) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  int j = perm[i];
  for (int a = 0; a < dim; ++a) dst[a * n + i] = src[a * n + j];
}

// occ[level_off + key] = 1 for every occupied cell of every level (from the leaf keys).
__global__ void k_mark_cells(const uint32_t* __restrict__ pkey, int64_t n, int dim, int height,
                             const int64_t* __restrict__ dense_off, int* __restrict__ occ) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint32_t k = pkey[i];
  if (i > 0 && pkey[i - 1] == k) return;  // one writer per leaf cell
  for (int l = height - 1; l >= 0; --l) {
    occ[dense_off[l] + k] = 1;
    k >>= dim;
  }
}

// After the exclusive scan: dense = occupied ? (scan - base[level]) : -1; keys[scan] = key.
__global__ void k_finish_levels(const int* __restrict__ occ, const int* __restrict__ scan,
                                int64_t total, int height, const int64_t* __restrict__ dense_off,
                                int* __restrict__ dense, uint32_t* __restrict__ keys) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  int l = 0;
  while (l + 1 < height && i >= dense_off[l + 1]) ++l;
  if (occ[i]) {
    int base = scan[dense_off[l]];
    dense[i] = scan[i] - base;
    keys[scan[i]] = static_cast<uint32_t>(i - dense_off[l]);
  } else {
    dense[i] = -1;
  }
}

// base[l] = first compact cell id of level l, [height] = number of cells, [height + 1] = "a point outside the box".
__global__ void k_level_summary(const int* __restrict__ scan, const int* __restrict__ occ,
                                const int64_t* __restrict__ off, int height, const int* __restrict__ outside,
                                int* __restrict__ out) {
  const int l = threadIdx.x;
  if (l < height) out[l] = scan[off[l]];
  if (l == height) {
    const int64_t total = off[height];
    out[height] = scan[total - 1] + occ[total - 1];
    out[height + 1] = *outside;
  }
}

__global__ void k_leaf_start(const uint32_t* __restrict__ pkey, int64_t n, const int* __restrict__ dense_leaf,
                             int n_leaf, int* __restrict__ leaf_start) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i > n) return;
  if (i == n) {
    leaf_start[n_leaf] = static_cast<int>(n);
    return;
  }
  uint32_t k = pkey[i];
  if (i == 0 || pkey[i - 1] != k) leaf_start[dense_leaf[k]] = static_cast<int>(i);
}

}  // namespace

void Tree::build(int dim, int height, const Box& box, const double* pos_caller, int64_t n,
                 cudaStream_t stream, LaunchCounter& ctr) {
  PLT_REQUIRE(dim >= 1 && dim <= 3, "dim must be 1, 2 or 3");
  PLT_REQUIRE(height >= 2 && height <= 23 && dim * (height - 1) <= 30, "tree too deep");
  PLT_REQUIRE(n > 0 && n < (int64_t{1} << 31), "point count out of range");
  dim_ = dim;
  n_ = n;
  const int leaf = height - 1;
  const int threads = 256;
  const int blocks = ceil_div(n, threads);

  // 1. keys + radix sort (key, caller index).
  key_tmp_.alloc(n, stream);
  idx_tmp_.alloc(n, stream);
  pkey_.alloc(n, stream);
  perm_.alloc(n, stream);
  // scan_[total] (one past the exclusive scan) carries the "point outside the root box" flag.
  dense_off_.assign(height + 1, 0);
  for (int l = 0; l < height; ++l) dense_off_[l + 1] = dense_off_[l] + (int64_t{1} << (dim * l));
  const int64_t total = dense_off_[height];
  scan_.alloc(total + 1, stream);
  int* outside = scan_.get() + total;
  PLT_CUDA(cudaMemsetAsync(outside, 0, sizeof(int), stream));
  if (dim == 1) PLT_LAUNCH(ctr, k_point_keys<1>, blocks, threads, 0, stream, pos_caller, n, box, leaf, key_tmp_.get(), idx_tmp_.get(), outside);
  if (dim == 2) PLT_LAUNCH(ctr, k_point_keys<2>, blocks, threads, 0, stream, pos_caller, n, box, leaf, key_tmp_.get(), idx_tmp_.get(), outside);
  if (dim == 3) PLT_LAUNCH(ctr, k_point_keys<3>, blocks, threads, 0, stream, pos_caller, n, box, leaf, key_tmp_.get(), idx_tmp_.get(), outside);
  size_t tmp_bytes = 0;
  const int end_bit = std::max(1, dim * leaf);
  PLT_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_tmp_.get(), pkey_.get(), idx_tmp_.get(),
                                           perm_.get(), static_cast<int>(n), 0, end_bit, stream));
  tmp_.alloc(tmp_bytes, stream);
  PLT_CUDA(cub::DeviceRadixSort::SortPairs(tmp_.get(), tmp_bytes, key_tmp_.get(), pkey_.get(), idx_tmp_.get(),
                                           perm_.get(), static_cast<int>(n), 0, end_bit, stream));
  ctr.n += 3;  // CUB onesweep: histogram + sweep passes (counted as library launches)

  // 2. sorted positions.
  pos_.alloc(static_cast<size_t>(dim) * n, stream);
  PLT_LAUNCH(ctr, k_gather_pos, blocks, threads, 0, stream, pos_caller, perm_.get(), n, dim, pos_.get());

  // 3. occupancy of every level, one scan over the concatenation.
  DevBuf<int64_t>& d_off = d_off_;
  d_off.alloc(height + 1, stream);
  PLT_CUDA(cudaMemcpyAsync(d_off.get(), dense_off_.data(), (height + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, stream));
  DevBuf<int>& occ = occ_;
  DevBuf<int>& scan = scan_;
  occ.alloc(total, stream);
  occ.zero(stream);
  PLT_LAUNCH(ctr, k_mark_cells, blocks, threads, 0, stream, pkey_.get(), n, dim, height, d_off.get(), occ.get());
  size_t scan_bytes = 0;
  PLT_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, occ.get(), scan.get(), static_cast<int>(total), stream));
  if (scan_bytes > tmp_.size()) tmp_.alloc(scan_bytes, stream);
  PLT_CUDA(cub::DeviceScan::ExclusiveSum(tmp_.get(), scan_bytes, occ.get(), scan.get(), static_cast<int>(total), stream));
  ctr.n += 1;

  // Cell counts per level: gathered on the device, ONE small copy into pinned memory; the only sync of the build.
  summary_.alloc(height + 2, stream);
  PLT_LAUNCH(ctr, k_level_summary, 1, 32, 0, stream, scan.get(), occ.get(), d_off.get(), height, outside,
             summary_.get());
  if (!h_summary_) {
    int* p = nullptr;
    PLT_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&p), sizeof(int) * 32, cudaHostAllocDefault));
    h_summary_.reset(p);
  }
  PLT_CUDA(cudaMemcpyAsync(h_summary_.get(), summary_.get(), sizeof(int) * (height + 2), cudaMemcpyDeviceToHost, stream));
  PLT_CUDA(cudaStreamSynchronize(stream));
  std::vector<int> base(h_summary_.get(), h_summary_.get() + height + 1);
  if (h_summary_.get()[height + 1]) {
    height_ = 0;
    throw Error(PLT_ERR_INVALID, "a point lies outside the bounding box the evaluator was created with");
  }
  total_cells_ = base[height];
  cell_off_.assign(height + 1, 0);
  n_cells_.assign(height, 0);
  for (int l = 0; l < height; ++l) {
    cell_off_[l] = base[l];
    n_cells_[l] = base[l + 1] - base[l];
  }
  cell_off_[height] = total_cells_;

  dense_.alloc(total, stream);
  keys_.alloc(total_cells_, stream);
  PLT_LAUNCH(ctr, k_finish_levels, ceil_div(total, threads), threads, 0, stream, occ.get(), scan.get(), total,
             height, d_off.get(), dense_.get(), keys_.get());

  // 4. leaf point ranges.
  leaf_start_.alloc(n_cells_[leaf] + 1, stream);
  PLT_LAUNCH(ctr, k_leaf_start, ceil_div(n + 1, threads), threads, 0, stream, pkey_.get(), n,
             dense_.get() + dense_off_[leaf], n_cells_[leaf], leaf_start_.get());
  height_ = height;
}

TreeView Tree::view() const {
  TreeView v{};
  v.dim = dim_;
  v.height = height_;
  v.n = n_;
  v.pos = pos_.get();
  v.perm = perm_.get();
  v.leaf_start = leaf_start_.get();
  v.dense = dense_.get();
  v.keys = keys_.get();
  for (int l = 0; l < height_; ++l) {
    v.dense_off[l] = dense_off_[l];
    v.cell_off[l] = cell_off_[l];
    v.n_cells[l] = n_cells_[l];
  }
  return v;
}

}  // namespace plt
