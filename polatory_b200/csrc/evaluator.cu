// Host-side evaluator: the state machine behind the C ABI (include/polatory_b200.h).
//
// Mirrors FmmGenericEvaluator<Kernel>::Impl (src/fmm/fmm_evaluator.hpp:32-293) and
// FmmGenericSymmetricEvaluator<Kernel>::Impl (src/fmm/fmm_symmetric_evaluator.hpp:31-275):
// same setters, same brute-force thresholds, same tree-height rule, same accuracy ->
// (order, d) policy; everything below that line is device-resident and new.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <numeric>
#include <random>
#include <string>
#include <tuple>
#include <vector>

#include <cub/cub.cuh>

#include "common.cuh"
#include "direct.cuh"
#include "fmm_ops.cuh"
#include "interp.hpp"
#include "plan.cuh"
#include "rbf_host.hpp"
#include "tree.cuh"

namespace plt {
namespace {

constexpr int kClassic = -1;
constexpr int64_t kMaxSearchTargets = 10000;  // fmm_accuracy_estimator.hpp:71

__global__ void k_max_abs_diff(const double* __restrict__ a, const double* __restrict__ b, int64_t n,
                               unsigned long long* __restrict__ out) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  double d = 0.0;
  if (i < n) {
    d = fabs(a[i] - b[i]);
    if (!(d == d)) d = INFINITY;  // NaN -> inf so that it never passes the accuracy test
  }
  for (int o = 16; o > 0; o >>= 1) d = fmax(d, __shfl_xor_sync(0xffffffffu, d, o));
  if ((threadIdx.x & 31) == 0 && d > 0.0) atomicMax(out, static_cast<unsigned long long>(__double_as_longlong(d)));
}

__global__ void k_gather_soa(const double* __restrict__ src, int64_t n_src, const int* __restrict__ idx,
                             int64_t n, int dim, double* __restrict__ dst) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  int j = idx[i];
  for (int a = 0; a < dim; ++a) dst[a * n + i] = src[a * n_src + j];
}

__global__ void k_add(const double* __restrict__ a, int64_t n, double* __restrict__ out) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) out[i] += a[i];
}

// Per-level compact-cell ranges covering the leaves [leaf_lo, leaf_hi) of a tree, and the
// matching slot ranges of the plan's active-parent lists (children level l <-> parents l-1).
__global__ void k_level_ranges(TreeView tr, PlanView pv, int leaf_lo, int leaf_hi, int* __restrict__ lo,
                               int* __restrict__ hi, int* __restrict__ slot_lo, int* __restrict__ slot_hi) {
  int l = threadIdx.x;
  if (l >= tr.height) return;
  const int leaf = tr.height - 1;
  if (leaf_hi <= leaf_lo) {
    lo[l] = hi[l] = 0;
    slot_lo[l] = slot_hi[l] = 0;
    return;
  }
  uint32_t k0 = tr.keys[tr.cell_off[leaf] + leaf_lo] >> (tr.dim * (leaf - l));
  uint32_t k1 = tr.keys[tr.cell_off[leaf] + leaf_hi - 1] >> (tr.dim * (leaf - l));
  lo[l] = tr.dense[tr.dense_off[l] + k0];
  hi[l] = tr.dense[tr.dense_off[l] + k1] + 1;
  slot_lo[l] = slot_hi[l] = 0;
  if (l >= 2) {
    // parents of level l-1 in [plo, phi): slots of level l whose parent id falls in that range
    const int plo = tr.dense[tr.dense_off[l - 1] + (k0 >> tr.dim)];
    const int phi = tr.dense[tr.dense_off[l - 1] + (k1 >> tr.dim)] + 1;
    const int b = pv.level_begin[l], e = pv.level_begin[l + 1];
    int x = b, y = e;
    while (x < y) {
      int m = (x + y) >> 1;
      if (pv.active[m] < plo) x = m + 1; else y = m;
    }
    slot_lo[l] = x;
    y = e;
    while (x < y) {
      int m = (x + y) >> 1;
      if (pv.active[m] < phi) x = m + 1; else y = m;
    }
    slot_hi[l] = x;
  }
}

// First leaf whose point range starts at or after `point` (leaf-granular shard boundary).
__global__ void k_find_leaf(const int* __restrict__ leaf_start, int n_leaf, int point, int* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_leaf) return;
  int s = leaf_start[i];
  int prev = i == 0 ? -1 : leaf_start[i - 1];
  if (s >= point && prev < point) *out = i;
}

// ---- partitioned upward pass (multi-GPU) --------------------------------------------------
// need[key] = 1 for every cell of level `cut` (by Morton key) that lies within the 3^dim neighbourhood of a
// target cell of that level: the source cells below it appear in the M2L / P2P lists of this rank's targets.
template <int DIM>
__global__ void k_need_mask(TreeView trg, int cut, unsigned char* __restrict__ need) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= trg.n_cells[cut]) return;
  int c[DIM];
  morton_decode<DIM>(trg.keys[trg.cell_off[cut] + i], c);
  const int nside = 1 << cut;
  constexpr int NN = DIM == 1 ? 3 : (DIM == 2 ? 9 : 27);
  for (int e = 0; e < NN; ++e) {
    int q[DIM], r = e;
    bool ok = true;
#pragma unroll
    for (int a = DIM - 1; a >= 0; --a) {
      q[a] = c[a] + (r % 3) - 1;
      r /= 3;
      ok = ok && q[a] >= 0 && q[a] < nside;
    }
    if (ok) need[morton_encode<DIM>(q)] = 1;
  }
}

// The same from the raw target points (before their tree exists, so that the partitioned upward pass can run while the
// target tree is being built): occ[key] = 1 for the level-`cut` cell of every point -- the leaf cell of tree.cu's
// k_point_keys, shifted up -- then the dilation by one cell.
template <int DIM>
__global__ void k_cut_cells_of_points(const double* __restrict__ pos, int64_t n, Box box, int leaf, int cut,
                                      unsigned char* __restrict__ occ) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int nside = 1 << leaf;
  const double inv_w = static_cast<double>(nside) / box.width;
  int c[DIM];
#pragma unroll
  for (int a = 0; a < DIM; ++a) {
    const double x = (pos[a * n + i] - (box.center[a] - 0.5 * box.width)) * inv_w;
    c[a] = min(max(static_cast<int>(floor(x)), 0), nside - 1);
  }
  const uint32_t key = morton_encode<DIM>(c) >> (DIM * (leaf - cut));
  if (!occ[key]) occ[key] = 1;
}
template <int DIM>
__global__ void k_dilate_need(const unsigned char* __restrict__ occ, int cut, unsigned char* __restrict__ need) {
  const uint32_t key = blockIdx.x * blockDim.x + threadIdx.x;
  if (key >= (1u << (DIM * cut)) || !occ[key]) return;
  int c[DIM];
  morton_decode<DIM>(key, c);
  const int nside = 1 << cut;
  constexpr int NN = DIM == 1 ? 3 : (DIM == 2 ? 9 : 27);
  for (int e = 0; e < NN; ++e) {
    int q[DIM], r = e;
    bool ok = true;
#pragma unroll
    for (int a = DIM - 1; a >= 0; --a) {
      q[a] = c[a] + (r % 3) - 1;
      r /= 3;
      ok = ok && q[a] >= 0 && q[a] < nside;
    }
    if (ok) need[morton_encode<DIM>(q)] = 1;
  }
}

// Work flags of every source cell (tree.cuh: kCellFlagM / kCellFlagMhat) for rank `rank` of a partition of the
// level-`cut` key space into the ranges [key_begin[r], key_begin[r + 1]):
//   levels <  cut : M by M2M from the all-gathered level `cut` and Mhat on every rank (a handful of cells)
//   level  == cut : M if owned or needed (the all-gather fills in the rest), Mhat on every rank
//   levels >  cut : M if the level-`cut` ancestor is owned or needed, Mhat if it is needed
__global__ void k_cell_flags(TreeView src, int cut, const unsigned char* __restrict__ need, uint32_t own_lo,
                             uint32_t own_hi, unsigned char* __restrict__ flags) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= src.cell_off[src.height - 1] + src.n_cells[src.height - 1]) return;
  int l = 0;
  while (l + 1 < src.height && g >= src.cell_off[l + 1]) ++l;
  unsigned char f;
  if (l < cut) {
    f = kCellFlagM | kCellFlagMhat;
  } else {
    const uint32_t anc = src.keys[g] >> (src.dim * (l - cut));
    const bool own = anc >= own_lo && anc < own_hi, nd = need[anc] != 0;
    f = (own || nd ? kCellFlagM : 0) | (l == cut || nd ? kCellFlagMhat : 0);
  }
  flags[g] = f;
}

// out[r] = first compact cell of `level` whose key is >= bounds[r] (n_cells if none), r < n.
__global__ void k_lower_bound_keys(TreeView tr, int level, const uint32_t* __restrict__ bounds, int n,
                                   int shift, int* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const uint32_t* keys = tr.keys + tr.cell_off[level];
  const unsigned long long key = static_cast<unsigned long long>(bounds[r]) << shift;
  int lo = 0, hi = tr.n_cells[level];
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (keys[mid] < key) lo = mid + 1; else hi = mid;
  }
  out[r] = lo;
}

struct PhaseTimer {
  struct Rec {
    const char* name;
    cudaEvent_t e0, e1;
  };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  size_t used = 0;
  cudaEvent_t get() {
    if (used == pool.size()) {
      cudaEvent_t e;
      PLT_CUDA(cudaEventCreate(&e));
      pool.push_back(e);
    }
    return pool[used++];
  }
  void reset() {
    recs.clear();
    used = 0;
  }
  void begin(const char* name, cudaStream_t s) {
    Rec r{name, get(), get()};
    PLT_CUDA(cudaEventRecord(r.e0, s));
    recs.push_back(r);
  }
  void end(cudaStream_t s) { PLT_CUDA(cudaEventRecord(recs.back().e1, s)); }
  ~PhaseTimer() {
    for (auto e : pool) cudaEventDestroy(e);
  }
};

struct ConfigKey {
  int height, order, d;
  bool operator<(const ConfigKey& o) const { return std::tie(height, order, d) < std::tie(o.height, o.order, o.d); }
};

// Block spectra of the parent-block M2L and the children levels they were built for (bit l).
struct BlkSpectra {
  DevBuf<double2> buf;
  unsigned levels = 0;
};

// Device-resident interpolator for one (height, order, d): tables + per-level M2L operators.
// Counterpart of scalfmm::interpolation::interpolator(kernel, order, tree_height, box_width, d)
// cached like the reference's LruCache<InterpolatorConfiguration, Interpolator>(2)
// (src/fmm/fmm_evaluator.hpp:249-252,292).
struct Interpolator {
  InterpTables host;
  DevBuf<double> beta, child;
  DevBuf<double2> tw;
  DevBuf<double2> khat;  // [levels 2..height-1][7^dim][kn][km][F]
  size_t khat_level_stride = 0;
  // Parent-block M2L (fmm_blk.cu): block operators [levels 2..height-1][27][kn][km][FB]; blk_ok = supported for this
  // (dim, order) and every tabulated value is finite (the block sums touch k at distance 0).
  DevBuf<double2> kblk, tw_blk;
  size_t kblk_level_stride = 0;
  bool blk_ok = false;
  InterpDev dev{};
  uint64_t last_use = 0;
};

}  // namespace
}  // namespace plt

using namespace plt;

struct plt_eval {
  int kind = 0, symmetric = 0, dim = 3, km = 1, kn = 1;
  int rbf_id = 0, rbf_part = 0;
  RbfHost rbf;
  std::vector<double> params;
  double aniso[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  double bbox_min[3] = {0, 0, 0}, bbox_max[3] = {0, 0, 0};
  Box box;
  double accuracy = std::numeric_limits<double>::infinity();
  int force_order = 0, force_d = kClassic, force_height = 0;
  bool force_direct = false;  // always take the brute-force branch (exact sums: the fit's residual sample)
  cudaStream_t stream = nullptr;
  // Side stream for the target tree + plan, concurrent with the upward pass (evaluate_device).  Declared before the
  // trees and the plan: their buffers are stream-ordered allocations of this stream and must be freed before it goes.
  struct SideStream {
    cudaStream_t s = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    ~SideStream() {
      if (fork) cudaEventDestroy(fork);
      if (join) cudaEventDestroy(join);
      if (s) cudaStreamDestroy(s);
    }
  } side;
  LaunchCounter ctr;
  std::string err;
  PhaseTimer timer;

  // Spheroidal split (src/fmm/spheroidal_evaluator.hpp:17-29).
  std::unique_ptr<plt_eval> direct_part, fast_part;

  int64_t n_src = 0, n_trg = 0;
  DevBuf<double> src_pos_c, trg_pos_c;  // caller order, transformed, SoA
  DevBuf<double> w_caller;              // weights as given [km * n_src]
  bool have_weights = false;

  Tree src_tree, trg_tree;
  Plan plan;           // interaction plan of (src_tree, target_tree()); rebuilt with either tree
  Arena arena;         // temporaries of one evaluate()
  DevBuf<double> stage;  // host -> device staging of caller points
  cudaStream_t copy_stream = nullptr;  // H2D of host targets, overlapped with the upward pass
  cudaEvent_t copy_ready = nullptr, copy_done = nullptr;
  // Slab streaming of host targets (evaluate_points): two staging buffers, the slab's transformed positions, a
  // stream for the D2H of finished slabs and three events per slab (copied in / staging buffer free / values ready).
  struct SlabPipe {
    cudaStream_t out_stream = nullptr;
    std::vector<cudaEvent_t> ev;
    DevBuf<double> stage[2], pos;
    const double* trg_pos = nullptr;  // positions the target tree is built from while a slab is evaluated
    ~SlabPipe() {
      for (auto e : ev) cudaEventDestroy(e);
      if (out_stream) cudaStreamDestroy(out_stream);
    }
  } slabs;
  bool multipole_dirty = true;
  int up_order = 0, up_d = 0;  // configuration the cached multipoles were built with
  DevBuf<double> wt_sorted;    // [km][n_src] folded, sorted
  bool wt_dirty = true;
  DevBuf<double> M;
  DevBuf<double2> Mhat;
  BlkSpectra Mblk;

  std::map<ConfigKey, std::unique_ptr<Interpolator>> interp_cache;
  uint64_t use_clock = 0;
  std::map<int, plt_config> best_config;  // per tree height (fmm_evaluator.hpp:187-197)
  plt_config config{0, 0, kClassic};

  int shard_rank = 0, shard_world = 1;

  // Partition of the level-`cut` Morton key space over the ranks of one node (plt_eval_set_partition): rank r
  // owns the keys [key_begin[r], key_begin[r + 1]).  Sources and weights are replicated; a rank computes the
  // multipoles of the cells it owns or needs below the cut, the level-`cut` expansions are all-gathered and
  // the upper levels are finished redundantly (they are a few hundred cells).
  struct Partition {
    bool on = false;
    int rank = 0, world = 1, cut = 0;
    std::vector<uint32_t> key_begin;
    plt_allgatherv_fn allgatherv = nullptr;
    void* ctx = nullptr;
  } part;
  DevBuf<unsigned char> need_mask, cell_flags;
  DevBuf<uint32_t> d_key_begin;
  std::vector<int> own_cells;   // [world + 1] compact level-`cut` cell ranges of the ranks (source tree)
  bool own_cells_valid = false, cell_flags_valid = false;

  plt_eval() = default;
  plt_eval(const plt_eval&) = delete;
  plt_eval& operator=(const plt_eval&) = delete;
  ~plt_eval() {
    if (copy_ready) cudaEventDestroy(copy_ready);
    if (copy_done) cudaEventDestroy(copy_done);
    if (copy_stream) cudaStreamDestroy(copy_stream);

  }

  // -------------------------------------------------------------------------------
  const Tree& target_tree() const { return symmetric ? src_tree : trg_tree; }
  int64_t targets() const { return symmetric ? n_src : n_trg; }
  const double* target_pos_c() const { return symmetric ? src_pos_c.get() : trg_pos_c.get(); }

  void init(int kind_, int symmetric_, int dim_, int rbf_id_, int part_, const double* params_, int n_params,
            const double* aniso_, const double* bmin, const double* bmax) {
    PLT_REQUIRE(dim_ >= 1 && dim_ <= 3, "dim must be 1, 2 or 3");
    PLT_REQUIRE(kind_ >= 0 && kind_ <= 3, "unknown kernel kind");
    PLT_REQUIRE(bmin && bmax, "bbox is required");
    kind = kind_;
    symmetric = symmetric_ ? 1 : 0;
    dim = dim_;
    rbf_id = rbf_id_;
    rbf_part = part_;
    km = (kind == KIND_F || kind == KIND_H) ? dim : 1;
    kn = (kind == KIND_FT || kind == KIND_H) ? dim : 1;
    // Symmetric evaluators exist only for K and H (fmm_symmetric_evaluator.hpp:74-78).
    PLT_REQUIRE(!symmetric || kind == KIND_K || kind == KIND_H, "symmetric evaluators are K or H only");
    try {
      rbf = make_rbf_const(rbf_id, part_, params_, n_params);
    } catch (const std::invalid_argument& e) {
      throw Error(PLT_ERR_INVALID, e.what());
    }
    // cov_spherical / cov_cubic have no Hessian: the reference still instantiates their Hessian
    // evaluators (src/fmm/make_fmm_evaluator.cpp CASE(CovCubic) / CASE(CovSpherical)) and only throws
    // from evaluate_hessian_isotropic (cov_spherical.hpp:53-55), i.e. when a pair is evaluated.
    params.assign(params_, params_ + n_params);
    for (int i = 0; i < dim * dim; ++i) aniso[i] = aniso_ ? aniso_[i] : (i / dim == i % dim ? 1.0 : 0.0);
    // rbf_base.hpp:73-75
    double det = dim == 1 ? aniso[0]
                 : dim == 2 ? aniso[0] * aniso[3] - aniso[1] * aniso[2]
                            : aniso[0] * (aniso[4] * aniso[8] - aniso[5] * aniso[7]) -
                                  aniso[1] * (aniso[3] * aniso[8] - aniso[5] * aniso[6]) +
                                  aniso[2] * (aniso[3] * aniso[7] - aniso[4] * aniso[6]);
    PLT_REQUIRE(det > 0.0, "aniso must have a positive determinant");
    for (int a = 0; a < dim; ++a) {
      bbox_min[a] = bmin[a];
      bbox_max[a] = bmax[a];
    }
    make_box();
    int ndev = 0;
    PLT_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev == 0) throw Error(PLT_ERR_CUDA, "no CUDA device");
    {
      // Keep freed blocks in the stream-ordered pool (default threshold 0 returns them to the
      // driver at every synchronisation, which turns each evaluate() into GBs of cudaMalloc).
      int dev = 0;
      PLT_CUDA(cudaGetDevice(&dev));
      cudaMemPool_t pool;
      PLT_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
      uint64_t keep = UINT64_MAX;
      PLT_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    if (rbf.spheroidal && part_ == PLT_PART_FULL) {
      direct_part = std::make_unique<plt_eval>();
      direct_part->init(kind, symmetric, dim, rbf_id, PLT_PART_DIRECT, params_, n_params, aniso_, bmin, bmax);
      fast_part = std::make_unique<plt_eval>();
      fast_part->init(kind, symmetric, dim, rbf_id, PLT_PART_FAST, params_, n_params, aniso_, bmin, bmax);
    }
  }

  // src/fmm/utility.hpp:18-33: bbox.transform(aniso) = bbox of the transformed corners.
  void make_box() {
    double lo[3], hi[3];
    for (int a = 0; a < dim; ++a) {
      lo[a] = std::numeric_limits<double>::infinity();
      hi[a] = -lo[a];
    }
    for (int c = 0; c < (1 << dim); ++c) {
      double p[3];
      for (int b = 0; b < dim; ++b) p[b] = ((c >> b) & 1) ? bbox_max[b] : bbox_min[b];
      for (int a = 0; a < dim; ++a) {
        double s = 0.0;
        for (int b = 0; b < dim; ++b) s += p[b] * aniso[a * dim + b];
        lo[a] = std::min(lo[a], s);
        hi[a] = std::max(hi[a], s);
      }
    }
    double width = 0.0;
    for (int a = 0; a < dim; ++a) width = std::max(width, hi[a] - lo[a]);
    width *= 1.01;
    if (width == 0.0) width = 1.0;
    box.width = width;
    for (int a = 0; a < dim; ++a) box.center[a] = lo[a] + 0.5 * (hi[a] - lo[a]);
  }

  static bool is_device_pointer(const void* p) {
    cudaPointerAttributes attr{};
    bool dev = cudaPointerGetAttributes(&attr, p) == cudaSuccess &&
               (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    return dev;
  }

  // Copy caller data (host or device pointer) in; host sources are fully consumed on return.
  void copy_in(void* dst, const void* src, size_t bytes) {
    if (!bytes) return;
    const bool dev = is_device_pointer(src);
    PLT_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream));
    if (!dev) PLT_CUDA(cudaStreamSynchronize(stream));
  }

  void ensure_copy_stream() {
    if (copy_stream) return;
    PLT_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
    PLT_CUDA(cudaEventCreateWithFlags(&copy_ready, cudaEventDisableTiming));
    PLT_CUDA(cudaEventCreateWithFlags(&copy_done, cudaEventDisableTiming));
  }

  // -------------------------------------------------------------------------------
  void set_points_impl(const double* pts, int64_t n, bool source) {
    PLT_REQUIRE(n >= 0 && (n == 0 || pts), "points");
    PLT_REQUIRE(n < (int64_t{1} << 31), "too many points");
    if (direct_part) {
      direct_part->set_points_impl(pts, n, source);
      fast_part->set_points_impl(pts, n, source);
    }
    const double* dev_pts = pts;
    if (n > 0 && !is_device_pointer(pts)) {
      stage.alloc(static_cast<size_t>(n) * dim, stream);
      dev_pts = stage.get();
      if (!source && can_prefetch_upward()) {
        // Host targets: their H2D copy (PCIe-bound, the longest single item of a host-buffer
        // evaluation) runs on a copy stream while this stream does the source-side upward pass,
        // which does not depend on the targets.
        ensure_copy_stream();
        PLT_CUDA(cudaEventRecord(copy_ready, stream));  // staging buffer allocated, previous use finished
        PLT_CUDA(cudaStreamWaitEvent(copy_stream, copy_ready, 0));
        prefetch_upward(n);                             // asynchronous launches on `stream`
        PLT_CUDA(cudaMemcpyAsync(stage.get(), pts, sizeof(double) * n * dim, cudaMemcpyHostToDevice, copy_stream));
        PLT_CUDA(cudaEventRecord(copy_done, copy_stream));
        PLT_CUDA(cudaStreamWaitEvent(stream, copy_done, 0));
      } else {
        PLT_CUDA(cudaMemcpyAsync(stage.get(), pts, sizeof(double) * n * dim, cudaMemcpyHostToDevice, stream));
      }
    }
    DevBuf<double>& pos = source ? src_pos_c : trg_pos_c;
    pos.alloc(static_cast<size_t>(n) * dim, stream);
    launch_transform_points(dim, aniso, dev_pts, n, pos.get(), stream, ctr);
    if (dev_pts != pts) PLT_CUDA(cudaStreamSynchronize(stream));  // the caller may reuse its host buffer
    plan.reset();
    if (source) {
      n_src = n;
      src_tree.reset();
      best_config.clear();   // fmm_evaluator.hpp:136-138
      multipole_dirty = true;
      wt_dirty = true;
      have_weights = false;
    } else {
      n_trg = n;
      trg_tree.reset();      // fmm_evaluator.hpp:155-156
    }
  }

  // The upward pass can be issued ahead of evaluate() when sources and weights are in place and
  // the FMM branch will be taken for `n_targets` targets.
  bool can_prefetch_upward() const {
    // (a partitioned upward pass depends on the targets: which source cells this rank needs)
    return !part.on && !direct_part && !symmetric && n_src > 0 && have_weights && !std::isfinite(rbf.support_radius);
  }
  void prefetch_upward(int64_t n_targets) {
    if (n_src * n_targets < int64_t{1024} * 1024 && force_height == 0) return;  // brute-force branch
    int height = fmm_tree_height(dim, std::max(n_src, n_targets));
    if (force_height > 0) height = force_height;
    if (height <= 2) return;
    if (!src_tree.built() || src_tree.height() != height) {
      src_tree.build(dim, height, box, src_pos_c.get(), n_src, stream, ctr);
      wt_dirty = true;
      multipole_dirty = true;
      plan.reset();
    }
    ensure_sorted_weights();
    const plt_config c = find_best_configuration(height);
    if (multipole_dirty || up_order != c.order || up_d != c.d) {
      Interpolator& ip = interpolator(height, c.order, c.d);
      upward(src_tree, wt_sorted.get(), ip, M, Mhat, Mblk, false);
      multipole_dirty = false;
      up_order = c.order;
      up_d = c.d;
    }
  }

  void set_weights(const double* w, int64_t len) {
    PLT_REQUIRE(len == km * n_src, "weights.rows() must be km * n_src_points");
    if (direct_part) {
      direct_part->set_weights(w, len);
      fast_part->set_weights(w, len);
    }
    w_caller.alloc(len, stream);
    copy_in(w_caller.get(), w, sizeof(double) * len);
    have_weights = true;
    wt_dirty = true;
    multipole_dirty = true;  // fmm_evaluator.hpp:181
  }

  void set_accuracy(double a) {
    accuracy = a;
    best_config.clear();  // fmm_evaluator.hpp:114-118
    if (direct_part) {
      direct_part->set_accuracy(a);
      fast_part->set_accuracy(a);
    }
  }

  // -------------------------------------------------------------------------------
  Interpolator& interpolator(int height, int order, int d) {
    ConfigKey key{height, order, d};
    auto it = interp_cache.find(key);
    if (it == interp_cache.end()) {
      // LRU(2), src/fmm/lru_cache.hpp.
      while (interp_cache.size() >= 2) {
        auto victim = interp_cache.begin();
        for (auto j = interp_cache.begin(); j != interp_cache.end(); ++j)
          if (j->second->last_use < victim->second->last_use) victim = j;
        interp_cache.erase(victim);
      }
      auto ip = std::make_unique<Interpolator>();
      PLT_REQUIRE(order >= 2 && order <= kMaxOrder, "interpolation order out of range");
      ip->host = make_interp_tables(order, d);
      const auto& h = ip->host;
      ip->beta.alloc(h.beta.size(), stream);
      ip->child.alloc(h.child.size(), stream);
      ip->tw.alloc(h.nf, stream);
      PLT_CUDA(cudaMemcpyAsync(ip->beta.get(), h.beta.data(), sizeof(double) * h.beta.size(), cudaMemcpyHostToDevice, stream));
      PLT_CUDA(cudaMemcpyAsync(ip->child.get(), h.child.data(), sizeof(double) * h.child.size(), cudaMemcpyHostToDevice, stream));
      PLT_CUDA(cudaMemcpyAsync(ip->tw.get(), h.tw.data(), sizeof(double2) * h.nf, cudaMemcpyHostToDevice, stream));
      PLT_CUDA(cudaStreamSynchronize(stream));  // host tables are about to go out of scope of the copy
      ip->dev = InterpDev{order, h.nf, ip->beta.get(), ip->child.get(), ip->tw.get(),
                          h.tw.data(), h.child.data(), h.beta.data()};
      ip->dev.polynomial = d < 0 || d >= order - 1;  // FH of degree order - 1 is the polynomial interpolant too
      const size_t F = freqs_per_cell(order, dim);
      ip->khat_level_stride = static_cast<size_t>(ipow(7, dim)) * kn * km * F;
      const int n_levels = std::max(0, height - 2);
      ip->khat.alloc(ip->khat_level_stride * n_levels, stream);
      ip->khat.zero(stream);
      for (int l = 2; l < height; ++l)
        launch_tabulate_m2l(kind, dim, rbf.k, box, l, ip->dev, ip->khat.get() + ip->khat_level_stride * (l - 2),
                            stream, ctr);
      if (blk_supported(dim, order) && n_levels > 0) {
        const int nb = blk_nf(order);
        std::vector<double2> twb(nb);
        const double pi = 3.14159265358979323846264338327950288;
        for (int i = 0; i < nb; ++i) twb[i] = make_double2(std::cos(2.0 * pi * i / nb), -std::sin(2.0 * pi * i / nb));
        ip->tw_blk.alloc(nb, stream);
        PLT_CUDA(cudaMemcpyAsync(ip->tw_blk.get(), twb.data(), sizeof(double2) * nb, cudaMemcpyHostToDevice, stream));
        ip->kblk_level_stride = static_cast<size_t>(27) * kn * km * blk_freqs(order);
        ip->kblk.alloc(ip->kblk_level_stride * n_levels, stream);
        ip->kblk.zero(stream);
        for (int l = 2; l < height; ++l)
          launch_tabulate_m2l_blk(kind, dim, rbf.k, box, l, order, ip->tw_blk.get(), ip->tw.get(),
                                  ip->kblk.get() + ip->kblk_level_stride * (l - 2),
                                  ip->khat.get() + ip->khat_level_stride * (l - 2), stream, ctr);
        DevBuf<int> bad;
        bad.alloc(1, stream);
        bad.zero(stream);
        launch_check_finite(reinterpret_cast<const double*>(ip->kblk.get()), 2 * ip->kblk.size(), bad.get(), stream, ctr);
        launch_check_finite(reinterpret_cast<const double*>(ip->khat.get()), 2 * ip->khat.size(), bad.get(), stream, ctr);
        int h_bad = 0;
        PLT_CUDA(cudaMemcpyAsync(&h_bad, bad.get(), sizeof(int), cudaMemcpyDeviceToHost, stream));
        PLT_CUDA(cudaStreamSynchronize(stream));
        ip->blk_ok = h_bad == 0;
        if (!ip->blk_ok) {  // k is singular at distance 0 for this kind: list path; drop the near operators again
          ip->kblk.release();
          ip->khat.zero(stream);
          for (int l = 2; l < height; ++l)
            launch_tabulate_m2l(kind, dim, rbf.k, box, l, ip->dev, ip->khat.get() + ip->khat_level_stride * (l - 2),
                                stream, ctr);
        }
      }
      it = interp_cache.emplace(key, std::move(ip)).first;
    }
    it->second->last_use = ++use_clock;
    return *it->second;
  }

  void ensure_sorted_weights() {
    if (!wt_dirty) return;
    PLT_REQUIRE(have_weights, "set_weights must be called before evaluate");
    wt_sorted.alloc(static_cast<size_t>(km) * n_src, stream);
    launch_prepare_weights(kind, dim, aniso, w_caller.get(), src_tree.perm(), n_src, wt_sorted.get(), stream, ctr);
    wt_dirty = false;
  }

  // P2M + M2M + multipole DFT (the reference's `fmm(src_tree, op, p2m | m2m)`,
  // src/fmm/fmm_evaluator.hpp:83-88).
  void upward(const Tree& st, const double* wt, Interpolator& ip, DevBuf<double>& M_, DevBuf<double2>& Mhat_,
              BlkSpectra& Mblk_, bool timed, bool partitioned = false) {
    const int order = ip.host.order;
    const size_t P = nodes_per_cell(order, dim), F = freqs_per_cell(order, dim);
    TreeView sv = st.view();
    M_.alloc(static_cast<size_t>(st.total_cells()) * km * P, stream);
    int cut = -1;
    if (partitioned) {
      ensure_partition_tables();
      sv.flags = cell_flags.get();
      cut = part_cut();
    }
    if (timed) timer.begin("p2m", stream);
    launch_p2m(dim, km, sv, box, ip.dev, wt, M_.get(), stream, ctr);
    if (timed) timer.end(stream);
    if (timed) timer.begin("m2m", stream);
    for (int l = st.height() - 2; l >= 2; --l) {
      if (l == cut - 1) {
        // levels >= cut are done for the cells this rank owns or needs: exchange the level-`cut` expansions
        // (every rank contributes the cells it owns), then finish the upper levels redundantly.
        if (timed) timer.end(stream);
        if (timed) timer.begin("allgather", stream);
        allgather_cut_level(st, M_.get(), P);
        if (timed) timer.end(stream);
        if (timed) timer.begin("m2m", stream);
      }
      launch_m2m(dim, km, sv, l, ip.dev, M_.get(), stream, ctr);
    }
    if (partitioned && cut >= 2 && cut - 1 < 2) {  // cut == 2: nothing above it, the exchange has not happened yet
      if (timed) timer.end(stream);
      if (timed) timer.begin("allgather", stream);
      allgather_cut_level(st, M_.get(), P);
      if (timed) timer.end(stream);
      if (timed) timer.begin("m2m", stream);
    }
    if (timed) timer.end(stream);
    size_t far_cells = 0;
    for (int l = 2; l < st.height(); ++l) far_cells += st.n_cells(l);
    Mhat_.alloc(far_cells * km * F, stream);
    if (timed) timer.begin("m2hat", stream);
    launch_m2hat(dim, km, sv, ip.dev, M_.get(), Mhat_.get(), stream, ctr);
    if (timed) timer.end(stream);
    // block spectra (parent-block M2L) of the parents of the levels that are dense enough, stored for the cells of
    // levels 1 .. deepest such parent level; an empty buffer selects the list path everywhere
    int blk_top = 0;  // deepest children level on the block path
    if (ip.blk_ok)
      for (int l = 2; l < st.height(); ++l)
        if (blk_level_dense(l, st.n_cells(l))) blk_top = l;
    size_t blk_cells = 0;
    for (int l = 1; l < blk_top; ++l) blk_cells += st.n_cells(l);
    const size_t blk_elems = blk_cells * km * blk_freqs(order);
    Mblk_.levels = 0;
    if (blk_top >= 2 && blk_elems * sizeof(double2) <= kBlkSpectraBudget) {
      Mblk_.buf.alloc(blk_elems, stream);
      if (timed) timer.begin("mblk", stream);
      for (int l = 2; l <= blk_top; ++l)  // runs of consecutive dense levels share a launch
        if (blk_level_dense(l, st.n_cells(l))) {
          int hi = l;
          while (hi + 1 <= blk_top && blk_level_dense(hi + 1, st.n_cells(hi + 1))) ++hi;
          launch_mblk(km, sv, order, l - 1, hi - 1, M_.get(), Mblk_.buf.get(), stream, ctr);
          for (int q = l; q <= hi; ++q) Mblk_.levels |= 1u << q;
          l = hi;
        }
      if (timed) timer.end(stream);
    } else {
      Mblk_.buf.alloc(0, stream);
    }
  }
  static constexpr size_t kBlkSpectraBudget = size_t{48} << 30;

  // In-place all-gather of M at the cut level: rank r's segment = the cells [own_cells[r], own_cells[r + 1]).
  void allgather_cut_level(const Tree& st, double* M_, size_t P) {
    const int cut = part_cut();
    std::vector<int64_t> off(part.world + 1);
    for (int r = 0; r <= part.world; ++r) off[r] = static_cast<int64_t>(own_cells[r]) * km * static_cast<int64_t>(P);
    double* base = M_ + static_cast<size_t>(st.view().cell_off[cut]) * km * P;
    PLT_REQUIRE(part.allgatherv != nullptr, "partitioned evaluation needs the all-gather callback");
    if (part.allgatherv(part.ctx, base, off.data(), part.world, stream) != 0)
      throw Error(PLT_ERR_INVALID, "all-gather callback failed");
    ++n_allgathers;
  }
  int64_t n_allgathers = 0;

  // M2L + L2L + L2P + P2P for the target leaves [leaf_lo, leaf_hi) -> vt (SoA [kn][n_trg], sorted).
  // vt must be zero on entry.
  void downward(const Tree& st, const double* wt, const DevBuf<double2>& Mhat_, const BlkSpectra& Mblk_,
                const Tree& tt, const Plan& pl, Interpolator& ip, double* vt, int leaf_lo, int leaf_hi, bool timed) {
    const int order = ip.host.order;
    const int height = tt.height(), leaf = height - 1;
    const size_t P = nodes_per_cell(order, dim), F = freqs_per_cell(order, dim);
    const int nc = 1 << dim, nn = dim == 1 ? 3 : (dim == 2 ? 9 : 27);
    TreeView sv = st.view(), tv = tt.view();
    const PlanView pv = pl.view();
    if (leaf_hi <= leaf_lo) return;

    // per-level compact ranges of the ancestors of the shard's leaves + active-slot ranges
    std::vector<int> lo(height, 0), hi(height, 0), slo(height, 0), shi(height, 0);
    if (leaf_lo == 0 && leaf_hi == tt.n_cells(leaf)) {
      for (int l = 0; l < height; ++l) {
        hi[l] = tt.n_cells(l);
        slo[l] = pv.level_begin[l];
        shi[l] = pv.level_begin[l + 1];
      }
    } else {
      int* d = arena.take<int>(4 * height);
      PLT_LAUNCH(ctr, k_level_ranges, 1, 32, 0, stream, tv, pv, leaf_lo, leaf_hi, d, d + height, d + 2 * height,
                 d + 3 * height);
      std::vector<int> h(4 * height);
      PLT_CUDA(cudaMemcpyAsync(h.data(), d, sizeof(int) * 4 * height, cudaMemcpyDeviceToHost, stream));
      PLT_CUDA(cudaStreamSynchronize(stream));
      for (int l = 0; l < height; ++l) {
        lo[l] = h[l];
        hi[l] = h[height + l];
        slo[l] = h[2 * height + l];
        shi[l] = h[3 * height + l];
      }
    }

    if (height > 2) {
      // Can the last level run fused (leaf expansions in shared memory only)?
      static const bool no_fused = getenv("PLT_DEBUG_NO_FUSED") != nullptr;  // A/B switch for parity bisection
      const bool fused = !no_fused && leaf_fused_supported(dim, order);
      // Locals of the levels that are materialised: 2 .. leaf-1 (fused) or 2 .. leaf.
      const size_t L_cells = fused ? tt.view().cell_off[leaf] : tt.total_cells();
      double* L = nullptr;
      if (leaf > 2 || !fused) {
        L = arena.take<double>(std::max<size_t>(L_cells, 1) * kn * P);
        PLT_CUDA(cudaMemsetAsync(L, 0, sizeof(double) * L_cells * kn * P, stream));
      }
      const bool use_blk = ip.blk_ok && Mblk_.levels != 0 && pv.n_groups > 0;
      const size_t FB = use_blk ? blk_freqs(order) : 0;
      const size_t per_parent = static_cast<size_t>(nc) * kn * F + kn * FB;
      const size_t budget = (static_cast<size_t>(use_blk ? 16 : 8) << 30) / sizeof(double2);
      const int chunk_parents = static_cast<int>(std::max<size_t>(1, budget / per_parent));
      int max_active = 0;
      for (int l = 2; l < height; ++l) max_active = std::max(max_active, shi[l] - slo[l]);
      const bool full_range_early = leaf_lo == 0 && leaf_hi == tt.n_cells(leaf);
      // (room for all levels side by side when that fits one chunk: the merged launches below)
      const int all_active = full_range_early ? shi[leaf] - slo[2] : 0;
      const size_t chunk_cap = static_cast<size_t>(std::min(chunk_parents, std::max(std::max(max_active, all_active), 1)));
      double2* Lhat = arena.take<double2>(chunk_cap * per_parent);
      double2* Lhat_blk = Lhat + chunk_cap * nc * kn * F;
      // compact M2L result of the leaf level (fused path), indexed by slot - level_begin[leaf]
      double* Lc = nullptr;
      const int n_leaf_active = pl.n_active(leaf);
      if (fused) Lc = arena.take<double>(static_cast<size_t>(std::max(n_leaf_active, 1)) * nc * kn * P);

      // The levels above the leaf level in ONE Hadamard launch (list path, all their slots at once): each of them is
      // a few hundred parents, latency-bound when launched one by one -- which is what a slab, a sampler batch or the
      // share of one rank of eight pays per evaluation.  Their spectra sit side by side in Lhat; the inverse
      // transforms and L2L then run level by level as before (same arithmetic per slot: bit-identical).
      const bool full_range = leaf_lo == 0 && leaf_hi == tt.n_cells(leaf);
      bool merged = false, merged_leaf = false;
      const int n_upper = leaf - 2;  // levels 2 .. leaf-1
      if (full_range && n_upper >= 1 && n_upper < 8 && m2l_hadamard_multi_level_supported()) {
        bool lists_only = true, leaf_lists = !(use_blk && ((Mblk_.levels >> leaf) & 1u));
        for (int l = 2; l < leaf; ++l) lists_only = lists_only && !(use_blk && ((Mblk_.levels >> l) & 1u));
        const int total_upper = shi[leaf - 1] - slo[2], total_all = shi[leaf] - slo[2];
        // a small evaluation (a sampler batch, the share of one rank of eight) takes the leaf level along: it fits the
        // Lhat chunk as well, and one launch instead of two is worth 0.07 ms of a 0.6 ms batch
        const bool with_leaf = leaf_lists && static_cast<size_t>(total_all) <= chunk_cap;
        if (lists_only && total_upper > 0 && static_cast<size_t>(total_upper) <= chunk_cap && (n_upper >= 2 || with_leaf)) {
          M2LArgs a{};
          a.trg = tv;
          a.level = 2;
          a.order = order;
          a.dim = dim;
          a.km = km;
          a.kn = kn;
          a.Mhat = Mhat_.get();
          a.Khat = ip.khat.get();
          a.khat_level_stride = ip.khat_level_stride;
          a.active = pv.active + slo[2];
          a.src_ids = pv.src_ids + static_cast<size_t>(slo[2]) * nn * nc;
          a.trg_mask = pv.trg_mask + slo[2];
          a.n_active = with_leaf ? total_all : total_upper;
          a.Lhat = Lhat;
          a.n_lvls = n_upper + (with_leaf ? 1 : 0);
          for (int i = 0; i < a.n_lvls; ++i) a.lvl_slot_end[i] = shi[2 + i] - slo[2];
          if (timed) timer.begin("m2l_hadamard", stream);
          launch_m2l_hadamard(a, stream, ctr);
          if (timed) timer.end(stream);
          // inverse transforms of the upper levels in one launch as well (a slot knows its level); the leaf level's go
          // to the compact buffer of the fused leaf pass and keep their own launch
          a.n_active = total_upper;
          a.n_lvls = n_upper;
          a.L = L;
          if (timed) timer.begin("m2l_idft", stream);
          launch_m2l_idft(a, ip.dev, stream, ctr);
          if (timed) timer.end(stream);
          merged = true;
          merged_leaf = with_leaf;
        }
      }

      for (int l = 2; l < height; ++l) {
        const int n_active = shi[l] - slo[l];
        const bool compact_out = fused && l == leaf;
        const bool had_done = merged && l < leaf;  // this level's spectra and locals are already done
        for (int c0 = 0; c0 < n_active; c0 += chunk_parents) {
          const int ncnk = std::min(chunk_parents, n_active - c0);
          const size_t s0 = static_cast<size_t>(slo[l]) + c0;
          M2LArgs a{};
          a.trg = tv;
          a.level = l;
          a.order = order;
          a.dim = dim;
          a.km = km;
          a.kn = kn;
          a.Mhat = Mhat_.get();
          a.Khat = ip.khat.get() + ip.khat_level_stride * (l - 2);
          a.active = pv.active + s0;
          a.src_ids = pv.src_ids + s0 * nn * nc;
          a.trg_mask = pv.trg_mask + s0;
          a.n_active = ncnk;
          a.Lhat = had_done ? Lhat + static_cast<size_t>(slo[l] - slo[2]) * nc * kn * F : Lhat;
          a.L = compact_out ? nullptr : L;
          a.Lc = compact_out ? Lc + (s0 - pv.level_begin[leaf]) * nc * kn * P : nullptr;
          if (use_blk && ((Mblk_.levels >> l) & 1u)) {
            // block sums over the neighbouring parents, minus the adjacent child pairs they contain (fmm_ops.cuh)
            a.Mblk = Mblk_.buf.get();
            a.Kblk = ip.kblk.get() + ip.kblk_level_stride * (l - 2);
            a.grp_first = pv.grp_first;
            a.grp_slot = pv.grp_slot;
            a.grp_src = pv.grp_src;
            a.grp_lo = pv.grp_level_begin[l];
            a.grp_hi = pv.grp_level_begin[l + 1];
            a.slot0 = static_cast<int>(s0);
            a.Lhat_blk = Lhat_blk;
            if (timed) timer.begin("m2l_blk", stream);
            launch_m2l_blk_hadamard(a, stream, ctr);
            if (timed) timer.end(stream);
            if (timed) timer.begin("m2l_near", stream);
            launch_m2l_hadamard_near(a, stream, ctr);
            if (timed) timer.end(stream);
            if (timed) timer.begin("m2l_idft", stream);
            launch_m2l_idft(a, ip.dev, stream, ctr);
            if (timed) timer.end(stream);
            if (timed) timer.begin("m2l_blk_idft", stream);
            launch_m2l_blk_idft(a, stream, ctr);
            if (timed) timer.end(stream);
            continue;
          }
          if (had_done) continue;  // spectra and inverse transforms of this level were done with the others
          if (merged_leaf && l == leaf) {
            a.Lhat = Lhat + static_cast<size_t>(slo[l] - slo[2]) * nc * kn * F;  // spectra already there
          } else {
            if (timed) timer.begin("m2l_hadamard", stream);
            launch_m2l_hadamard(a, stream, ctr);
            if (timed) timer.end(stream);
          }
          if (timed) timer.begin("m2l_idft", stream);
          launch_m2l_idft(a, ip.dev, stream, ctr);
          if (timed) timer.end(stream);
        }
        if (l > 2 && !compact_out) {
          if (timed) timer.begin("l2l", stream);
          launch_l2l(dim, kn, tv, l, ip.dev, L, lo[l], hi[l], lo[l - 1], hi[l - 1], stream, ctr);
          if (timed) timer.end(stream);
        }
      }
      if (timed) timer.begin(fused ? "l2l_l2p_leaf" : "l2p", stream);
      bool done = false;
      if (fused)
        done = launch_l2l_l2p_leaf(dim, kn, tv, box, ip.dev, leaf > 2 ? L : nullptr, Lc, pv.leaf_meta, vt, leaf_lo,
                                   leaf_hi, lo[leaf - 1], hi[leaf - 1], stream, ctr);
      if (!done) {
        PLT_REQUIRE(!fused, "fused leaf pass rejected an order it was planned for");
        launch_l2p(dim, kn, tv, box, ip.dev, L, vt, leaf_lo, leaf_hi, stream, ctr);
      }
      if (timed) timer.end(stream);
    }
    if (timed) timer.begin("p2p", stream);
    launch_p2p(kind, dim, rbf.k, sv, wt, tv, vt, pv.p2p_leaves, pv.n_p2p, leaf_lo, leaf_hi, stream, ctr);
    if (timed) timer.end(stream);
  }

  // Brute force in caller order: the tree_height == 0 branch.
  void brute_force(const double* spos, const double* w_c, int64_t ns, const double* tpos, int64_t nt,
                   double* out_caller) {
    double* wt = arena.take<double>(static_cast<size_t>(km) * ns);
    double* vt = arena.take<double>(static_cast<size_t>(kn) * nt);
    launch_prepare_weights(kind, dim, aniso, w_c, nullptr, ns, wt, stream, ctr);
    DirectArgs a{};
    a.k = rbf.k;
    a.spos = spos;
    a.swt = wt;
    a.ns = ns;
    a.tpos = tpos;
    a.nt = nt;
    a.out = vt;
    a.n_chunks = (ns > 0 && nt > 0) ? direct_plan_chunks(ns, nt) : 1;
    a.partial = a.n_chunks > 1 ? arena.take<double>(static_cast<size_t>(a.n_chunks) * kn * nt) : nullptr;
    a.symmetric = 0;
    launch_direct(kind, dim, a, stream, ctr);
    launch_finish_outputs(kind, dim, aniso, vt, nullptr, nt, 0, nt, out_caller, stream, ctr);
  }

  // src/fmm/fmm_accuracy_estimator.hpp:74-121.
  plt_config find_best_configuration(int height) {
    if (force_order > 0) return {height, force_order, force_d};
    auto it = best_config.find(height);
    if (it != best_config.end()) return it->second;
    plt_config c{height, 0, kClassic};
    if (std::isinf(accuracy) && accuracy > 0) {
      c = {height, 6, kClassic};
    } else if (accuracy == 0.0) {
      c = {height, 12, 8};
    } else {
      c = search_configuration(height);
    }
    best_config[height] = c;
    return c;
  }

  plt_config search_configuration(int height) {
    // "Errors at the data points are larger than those at randomly distributed points":
    // sample min(n_src, 10000) source points (of the Morton-sorted container) as targets.
    const int64_t nt = std::min<int64_t>(n_src, kMaxSearchTargets);
    std::mt19937 gen;
    std::vector<int> idx(n_src);
    std::iota(idx.begin(), idx.end(), 0);
    std::shuffle(idx.begin(), idx.end(), gen);
    const auto outer_mark = arena.mark();
    int* d_idx = arena.take<int>(nt);
    PLT_CUDA(cudaMemcpyAsync(d_idx, idx.data(), sizeof(int) * nt, cudaMemcpyHostToDevice, stream));
    double* tpos = arena.take<double>(static_cast<size_t>(dim) * nt);
    PLT_LAUNCH(ctr, k_gather_soa, ceil_div(nt, 256), 256, 0, stream, src_tree.pos(), n_src, d_idx, nt, dim, tpos);
    PLT_CUDA(cudaStreamSynchronize(stream));
    ensure_sorted_weights();

    // exact: brute force over the sorted sources with the folded sorted weights
    double* exact_raw = arena.take<double>(static_cast<size_t>(kn) * nt);
    double* approx_raw = arena.take<double>(static_cast<size_t>(kn) * nt);
    double* exact = arena.take<double>(static_cast<size_t>(kn) * nt);
    double* approx = arena.take<double>(static_cast<size_t>(kn) * nt);
    DirectArgs a{};
    a.k = rbf.k;
    a.spos = src_tree.pos();
    a.swt = wt_sorted.get();
    a.ns = n_src;
    a.tpos = tpos;
    a.nt = nt;
    a.out = exact_raw;
    a.n_chunks = direct_plan_chunks(n_src, nt);
    a.partial = a.n_chunks > 1 ? arena.take<double>(static_cast<size_t>(a.n_chunks) * kn * nt) : nullptr;
    launch_direct(kind, dim, a, stream, ctr);
    launch_finish_outputs(kind, dim, aniso, exact_raw, nullptr, nt, 0, nt, exact, stream, ctr);

    Tree sample_tree;
    sample_tree.build(dim, height, box, tpos, nt, stream, ctr);
    Plan sample_plan;
    sample_plan.build(src_tree, sample_tree, stream, ctr);
    unsigned long long* d_err = arena.take<unsigned long long>(1);
    DevBuf<double> M_;
    DevBuf<double2> Mhat_;
    BlkSpectra Mblk_;
    plt_config found{0, 0, kClassic};
    for (int order = 8; order <= 20 && !found.order; order += 2) {
      const int min_d = order >= 12 ? 7 : kClassic;
      const int max_d = order >= 12 ? 9 : kClassic;
      for (int d = min_d; d <= max_d; ++d) {
        const auto inner_mark = arena.mark();
        Interpolator& ip = interpolator(height, order, d);
        upward(src_tree, wt_sorted.get(), ip, M_, Mhat_, Mblk_, false);
        PLT_CUDA(cudaMemsetAsync(approx_raw, 0, sizeof(double) * kn * nt, stream));
        downward(src_tree, wt_sorted.get(), Mhat_, Mblk_, sample_tree, sample_plan, ip, approx_raw, 0,
                 sample_tree.n_cells(height - 1), false);
        launch_finish_outputs(kind, dim, aniso, approx_raw, sample_tree.perm(), nt, 0, nt, approx, stream, ctr);
        PLT_CUDA(cudaMemsetAsync(d_err, 0, sizeof(unsigned long long), stream));
        PLT_LAUNCH(ctr, k_max_abs_diff, ceil_div(kn * nt, 256), 256, 0, stream, approx, exact, kn * nt, d_err);
        unsigned long long bits = 0;
        PLT_CUDA(cudaMemcpyAsync(&bits, d_err, sizeof(bits), cudaMemcpyDeviceToHost, stream));
        PLT_CUDA(cudaStreamSynchronize(stream));
        arena.rewind(inner_mark);
        double err_abs;
        std::memcpy(&err_abs, &bits, sizeof(double));
        if (err_abs <= accuracy) {
          found = {height, order, d};
          break;
        }
      }
    }
    PLT_CUDA(cudaStreamSynchronize(stream));
    arena.rewind(outer_mark);
    if (!found.order)
      throw Error(PLT_ERR_ACCURACY, "failed to construct an evaluator that meets the desired accuracy");
    return found;
  }

  // Compact-support kernels (cov_spherical, cov_cubic, spheroidal direct parts): the reference
  // uses a kd-tree radius search (src/fmm/direct_evaluator.hpp:41-68); here a uniform cell list
  // with cell width >= support radius, i.e. the P2P kernel alone on a tree of suitable height.
  int compact_height() const {
    const double r = rbf.support_radius;
    int level = 0;
    while (level < 10 && box.width / static_cast<double>(1 << (level + 1)) >= r) ++level;
    const int cap = fmm_tree_height(dim, std::max<int64_t>(std::max(n_src, targets()), 2)) + 1;
    return std::max(2, std::min(level + 1, cap));
  }

  void evaluate(double* out, int64_t len) {
    const int64_t nt = targets();
    PLT_REQUIRE(len == kn * nt, "output length must be kn * n_trg_points");
    PLT_REQUIRE(have_weights || n_src == 0, "set_weights must be called before evaluate");
    timer.reset();
    arena.reset();
    const bool out_is_device = out && is_device_pointer(out);
    double* dst = out;
    if (!out_is_device) dst = arena.take<double>(std::max<int64_t>(len, 1));
    evaluate_device(dst);
    if (!out_is_device) {
      if (len) PLT_CUDA(cudaMemcpyAsync(out, dst, sizeof(double) * len, cudaMemcpyDeviceToHost, stream));
      PLT_CUDA(cudaStreamSynchronize(stream));
    }
  }

  // Bulk evaluation at new points: set_target_points(points) + evaluate(out) as ONE call -- the sequence of
  // interpolation::Evaluator::evaluate(points) (include/polatory/interpolation/evaluator.hpp:83-87).  With HOST buffers
  // on the FMM branch the targets are streamed in caller-order slabs: slab i+1 is copied in on the copy stream and the
  // values of slab i-1 are copied out on a third stream while slab i is evaluated.  Every slab is evaluated against
  // the octree of the WHOLE problem (same height, hence the same cells, lists and kernels), so the values are
  // bit-identical to the one-shot evaluation; the multipoles are computed once (while slab 0 is on its way in).
  // On return the evaluator is in the state set_target_points(points) + evaluate() leave it in.
  static constexpr int64_t kSlabMinTargets = int64_t{1} << 20;
  int slab_count(const double* pts, int64_t n, const double* out) const {
    static const char* env = getenv("PLT_SLABS");  // A/B switch: 1 = off, k = k slabs
    if (!can_prefetch_upward() || force_direct || shard_world > 1) return 1;
    if (n < 2 * kSlabMinTargets || is_device_pointer(pts) || is_device_pointer(out)) return 1;
    const int height = force_height > 0 ? force_height : fmm_tree_height(dim, std::max(n_src, n));
    if (height <= 2) return 1;
    // Every slab pays ~1 ms of fixed cost (its own target tree, plan and per-level launches), measured on config #3
    // (profiles/r02_j_slabs.md): two slabs from 2M targets, three (half-size edge slabs) from 6M.
    const int k = env ? atoi(env) : (n >= 6 * kSlabMinTargets ? 3 : 2);
    return static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(k, n / kSlabMinTargets)));
  }

  void evaluate_points(const double* pts, int64_t n, double* out, int64_t len) {
    PLT_REQUIRE(!symmetric, "evaluate_points is not available on a symmetric evaluator");
    PLT_REQUIRE(n >= 0 && (n == 0 || pts), "points");
    PLT_REQUIRE(len == kn * n, "output length must be kn * n_trg_points");
    const int n_slabs = slab_count(pts, n, out);
    if (n_slabs < 2) {
      set_points_impl(pts, n, false);
      evaluate(out, len);
      return;
    }
    timer.reset();
    arena.reset();
    SlabPipe& sp = slabs;
    ensure_copy_stream();
    if (!sp.out_stream) PLT_CUDA(cudaStreamCreateWithFlags(&sp.out_stream, cudaStreamNonBlocking));
    while (sp.ev.size() < static_cast<size_t>(3 * n_slabs)) {
      cudaEvent_t e;
      PLT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      sp.ev.push_back(e);
    }
    // The first slab's copy-in and the last slab's copy-out are the exposed ends of the pipeline: edge slabs are
    // smaller than the inner ones (weights e, 1, ..., 1, e).
    static const double edge = getenv("PLT_SLAB_EDGE") ? atof(getenv("PLT_SLAB_EDGE")) : 0.5;
    std::vector<int64_t> bound(n_slabs + 1, 0);
    {
      const double e = n_slabs > 2 ? std::min(1.0, std::max(0.05, edge)) : 1.0;
      const double total = 2.0 * e + (n_slabs - 2);
      double acc = 0.0;
      for (int i = 0; i < n_slabs; ++i) {
        acc += (i == 0 || i == n_slabs - 1) ? e : 1.0;
        bound[i + 1] = std::min<int64_t>(n, static_cast<int64_t>(std::llround(acc / total * static_cast<double>(n))));
      }
      bound[n_slabs] = n;
    }
    auto first = [&](int i) { return bound[std::min(i, n_slabs)]; };
    int64_t per = 0;
    for (int i = 0; i < n_slabs; ++i) per = std::max(per, bound[i + 1] - bound[i]);
    sp.stage[0].alloc(static_cast<size_t>(per) * dim, stream);
    sp.stage[1].alloc(static_cast<size_t>(per) * dim, stream);
    sp.pos.alloc(static_cast<size_t>(per) * dim, stream);
    trg_pos_c.alloc(static_cast<size_t>(n) * dim, stream);
    double* dst = arena.take<double>(len);
    PLT_CUDA(cudaEventRecord(copy_ready, stream));  // buffers allocated, their previous uses finished
    PLT_CUDA(cudaStreamWaitEvent(copy_stream, copy_ready, 0));
    PLT_CUDA(cudaStreamWaitEvent(sp.out_stream, copy_ready, 0));
    auto copy_in = [&](int i) {
      const int64_t o = first(i), m = first(i + 1) - o;
      PLT_CUDA(cudaMemcpyAsync(sp.stage[i & 1].get(), pts + o * dim, sizeof(double) * m * dim, cudaMemcpyHostToDevice,
                               copy_stream));
      PLT_CUDA(cudaEventRecord(sp.ev[3 * i], copy_stream));
    };
    const int saved_height = force_height;
    const int height = saved_height > 0 ? saved_height : fmm_tree_height(dim, std::max(n_src, n));
    auto restore = [&] {
      force_height = saved_height;
      sp.trg_pos = nullptr;
      n_trg = n;            // as set_target_points(points) leaves it: positions in place, tree and plan stale
      trg_tree.reset();
      plan.reset();
    };
    try {
      copy_in(0);
      prefetch_upward(n);   // main stream: the upward pass does not depend on the targets
      copy_in(1);
      force_height = height;
      for (int i = 0; i < n_slabs; ++i) {
        const int64_t o = first(i), m = first(i + 1) - o;
        if (m == 0) break;
        PLT_CUDA(cudaStreamWaitEvent(stream, sp.ev[3 * i], 0));
        launch_transform_points(dim, aniso, sp.stage[i & 1].get(), m, sp.pos.get(), stream, ctr);
        launch_transform_points(dim, aniso, sp.stage[i & 1].get(), m, trg_pos_c.get() + o, stream, ctr, n);
        PLT_CUDA(cudaEventRecord(sp.ev[3 * i + 1], stream));
        if (i + 2 < n_slabs && first(i + 2) < n) {
          PLT_CUDA(cudaStreamWaitEvent(copy_stream, sp.ev[3 * i + 1], 0));
          copy_in(i + 2);
        }
        n_trg = m;
        sp.trg_pos = sp.pos.get();
        trg_tree.reset();
        plan.reset();
        const auto mark = arena.mark();
        evaluate_device(dst + kn * o);
        arena.rewind(mark);
        PLT_CUDA(cudaEventRecord(sp.ev[3 * i + 2], stream));
        PLT_CUDA(cudaStreamWaitEvent(sp.out_stream, sp.ev[3 * i + 2], 0));
        PLT_CUDA(cudaMemcpyAsync(out + kn * o, dst + kn * o, sizeof(double) * kn * m, cudaMemcpyDeviceToHost,
                                 sp.out_stream));
      }
    } catch (...) {
      restore();
      cudaStreamSynchronize(copy_stream);
      cudaStreamSynchronize(stream);
      cudaStreamSynchronize(sp.out_stream);
      throw;
    }
    restore();
    PLT_CUDA(cudaStreamSynchronize(copy_stream));
    PLT_CUDA(cudaStreamSynchronize(stream));
    PLT_CUDA(cudaStreamSynchronize(sp.out_stream));
  }

  // Leaf range [leaf_lo, leaf_hi) and sorted point range [p_lo, p_hi) of the current target shard.
  // Cached per (target tree build, rank, world): the steady-state sharded matvec has no extra sync.
  struct ShardBounds {
    int rank = -1, world = 0;
    int leaf_lo = 0, leaf_hi = 0, p_lo = 0, p_hi = 0;
    bool valid = false;
  } shard_cache;

  void shard_leaves(const Tree& tt, int& leaf_lo, int& leaf_hi, int& p_lo, int& p_hi) {
    const int n_leaf = tt.n_cells(tt.height() - 1);
    leaf_lo = 0;
    leaf_hi = n_leaf;
    p_lo = 0;
    p_hi = static_cast<int>(tt.n());
    if (part.on) {
      // a generic evaluator is handed this rank's targets only; the symmetric one (the matvec) evaluates the
      // leaves of its own key range
      if (!symmetric) return;
      if (!(shard_cache.valid && shard_cache.rank == -2)) {
        ensure_partition_tables();
        TreeView tv = tt.view();
        const int leaf = tt.height() - 1;
        int* d = arena.take<int>(2);
        PLT_LAUNCH(ctr, k_lower_bound_keys, 1, 32, 0, stream, tv, leaf, d_key_begin.get() + part.rank, 2,
                   dim * (leaf - part_cut()), d);
        int lh[2] = {0, 0}, ph[2] = {0, 0};
        PLT_CUDA(cudaMemcpyAsync(lh, d, sizeof(lh), cudaMemcpyDeviceToHost, stream));
        PLT_CUDA(cudaStreamSynchronize(stream));
        PLT_CUDA(cudaMemcpyAsync(&ph[0], tv.leaf_start + lh[0], sizeof(int), cudaMemcpyDeviceToHost, stream));
        PLT_CUDA(cudaMemcpyAsync(&ph[1], tv.leaf_start + lh[1], sizeof(int), cudaMemcpyDeviceToHost, stream));
        PLT_CUDA(cudaStreamSynchronize(stream));
        shard_cache = ShardBounds{-2, part.world, lh[0], lh[1], ph[0], ph[1], true};
      }
      leaf_lo = shard_cache.leaf_lo;
      leaf_hi = shard_cache.leaf_hi;
      p_lo = shard_cache.p_lo;
      p_hi = shard_cache.p_hi;
      return;
    }
    if (shard_world <= 1) return;
    if (!(shard_cache.valid && shard_cache.rank == shard_rank && shard_cache.world == shard_world)) {
      TreeView tv = tt.view();
      DevBuf<int> dbuf;
      dbuf.alloc(4, stream);
      int* d = dbuf.get();
      int bounds[4] = {0, n_leaf, 0, 0};
      PLT_CUDA(cudaMemcpyAsync(d, bounds, sizeof(bounds), cudaMemcpyHostToDevice, stream));
      const int64_t n = tt.n();
      const int p0 = static_cast<int>(n * shard_rank / shard_world);
      const int p1 = static_cast<int>(n * (shard_rank + 1) / shard_world);
      if (shard_rank > 0)
        PLT_LAUNCH(ctr, k_find_leaf, ceil_div(n_leaf + 1, 256), 256, 0, stream, tv.leaf_start, n_leaf, p0, d);
      if (shard_rank + 1 < shard_world)
        PLT_LAUNCH(ctr, k_find_leaf, ceil_div(n_leaf + 1, 256), 256, 0, stream, tv.leaf_start, n_leaf, p1, d + 1);
      PLT_CUDA(cudaMemcpyAsync(bounds, d, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
      PLT_CUDA(cudaStreamSynchronize(stream));
      PLT_CUDA(cudaMemcpyAsync(&bounds[2], tv.leaf_start + bounds[0], sizeof(int), cudaMemcpyDeviceToHost, stream));
      PLT_CUDA(cudaMemcpyAsync(&bounds[3], tv.leaf_start + bounds[1], sizeof(int), cudaMemcpyDeviceToHost, stream));
      PLT_CUDA(cudaStreamSynchronize(stream));
      shard_cache = ShardBounds{shard_rank, shard_world, bounds[0], bounds[1], bounds[2], bounds[3], true};
    }
    leaf_lo = shard_cache.leaf_lo;
    leaf_hi = shard_cache.leaf_hi;
    p_lo = shard_cache.p_lo;
    p_hi = shard_cache.p_hi;
  }

  // True when evaluate() takes the brute-force branch (no tree).
  bool brute_force_branch() const {
    const bool compact = std::isfinite(rbf.support_radius);
    const int64_t nt = targets();
    const bool small = symmetric ? n_src < 1024 : n_src * nt < int64_t{1024} * 1024;
    return force_direct || (small && !compact && force_height == 0) ||
           (compact && compact_height() <= 2 && shard_world == 1);
  }

  // Builds the source / target trees of the FMM (or compact) branch if they are not current.
  int planned_height() const {
    const bool compact = std::isfinite(rbf.support_radius);
    int height = compact ? compact_height()
                         : fmm_tree_height(dim, symmetric ? n_src : std::max(n_src, targets()));
    if (force_height > 0) height = force_height;
    return height;
  }
  void ensure_trees() {
    ensure_src_tree();
    ensure_trg_tree(stream);
  }
  void ensure_src_tree() {
    const int height = planned_height();
    if (!src_tree.built() || src_tree.height() != height) {
      src_tree.build(dim, height, box, src_pos_c.get(), n_src, stream, ctr);
      wt_dirty = true;
      multipole_dirty = true;
      plan.reset();
      if (symmetric) shard_cache.valid = false;
      own_cells_valid = cell_flags_valid = false;
    }
  }
  bool trg_tree_stale() const {
    return !symmetric && (!trg_tree.built() || trg_tree.height() != planned_height());
  }
  void ensure_trg_tree(cudaStream_t s) {
    const int height = planned_height();
    if (trg_tree_stale()) {
      trg_tree.build(dim, height, box, slabs.trg_pos ? slabs.trg_pos : trg_pos_c.get(), n_trg, s, ctr);
      plan.reset();
      shard_cache.valid = false;
      if (part.on) {  // another set of needed source cells: the cached multipoles may not cover it
        cell_flags_valid = false;
        multipole_dirty = true;
      }
    }
  }

  int part_cut() const { return std::min(part.cut, src_tree.height() - 1); }

  // Compact level-`cut` cell ranges owned by the ranks, and this rank's work flags.
  void ensure_partition_tables() {
    const int cut = part_cut();
    const TreeView sv = src_tree.view();
    if (!own_cells_valid) {
      d_key_begin.alloc(part.world + 1, stream);
      // keys were given at level part.cut; a shallower tree compares at its leaf level
      std::vector<uint32_t> kb(part.key_begin);
      for (auto& k : kb) k >>= dim * (part.cut - cut);
      kb[part.world] = 1u << (dim * cut);
      PLT_CUDA(cudaMemcpyAsync(d_key_begin.get(), kb.data(), sizeof(uint32_t) * (part.world + 1),
                               cudaMemcpyHostToDevice, stream));
      int* d = arena.take<int>(part.world + 1);
      PLT_LAUNCH(ctr, k_lower_bound_keys, 1, 64, 0, stream, sv, cut, d_key_begin.get(), part.world + 1, 0, d);
      own_cells.assign(part.world + 1, 0);
      PLT_CUDA(cudaMemcpyAsync(own_cells.data(), d, sizeof(int) * (part.world + 1), cudaMemcpyDeviceToHost, stream));
      PLT_CUDA(cudaStreamSynchronize(stream));
      own_cells[part.world] = src_tree.n_cells(cut);
      own_key_lo = kb[part.rank];
      own_key_hi = kb[part.rank + 1];
      own_cells_valid = true;
      cell_flags_valid = false;
    }
    if (!cell_flags_valid) {
      const size_t n_keys = size_t{1} << (dim * cut);
      need_mask.alloc(n_keys, stream);
      need_mask.zero(stream);
      if (!symmetric && !trg_tree.built()) {
        // the target tree is built concurrently (evaluate_device): take the cells from the points themselves
        unsigned char* occ = arena.take<unsigned char>(n_keys);
        PLT_CUDA(cudaMemsetAsync(occ, 0, n_keys, stream));
        const int leaf = src_tree.height() - 1;
        const int gp = ceil_div(n_trg, 256), gk = ceil_div(static_cast<int64_t>(n_keys), 128);
        if (dim == 1) PLT_LAUNCH(ctr, k_cut_cells_of_points<1>, gp, 256, 0, stream, trg_pos_c.get(), n_trg, box, leaf, cut, occ);
        if (dim == 2) PLT_LAUNCH(ctr, k_cut_cells_of_points<2>, gp, 256, 0, stream, trg_pos_c.get(), n_trg, box, leaf, cut, occ);
        if (dim == 3) PLT_LAUNCH(ctr, k_cut_cells_of_points<3>, gp, 256, 0, stream, trg_pos_c.get(), n_trg, box, leaf, cut, occ);
        if (dim == 1) PLT_LAUNCH(ctr, k_dilate_need<1>, gk, 128, 0, stream, occ, cut, need_mask.get());
        if (dim == 2) PLT_LAUNCH(ctr, k_dilate_need<2>, gk, 128, 0, stream, occ, cut, need_mask.get());
        if (dim == 3) PLT_LAUNCH(ctr, k_dilate_need<3>, gk, 128, 0, stream, occ, cut, need_mask.get());
      } else {
        const TreeView tv = target_tree().view();
        const int nt = tv.n_cells[cut];
        if (dim == 1) PLT_LAUNCH(ctr, k_need_mask<1>, ceil_div(nt, 128), 128, 0, stream, tv, cut, need_mask.get());
        if (dim == 2) PLT_LAUNCH(ctr, k_need_mask<2>, ceil_div(nt, 128), 128, 0, stream, tv, cut, need_mask.get());
        if (dim == 3) PLT_LAUNCH(ctr, k_need_mask<3>, ceil_div(nt, 128), 128, 0, stream, tv, cut, need_mask.get());
      }
      cell_flags.alloc(src_tree.total_cells(), stream);
      PLT_LAUNCH(ctr, k_cell_flags, ceil_div(src_tree.total_cells(), 256), 256, 0, stream, sv, cut, need_mask.get(),
                 own_key_lo, own_key_hi, cell_flags.get());
      cell_flags_valid = true;
    }
  }
  uint32_t own_key_lo = 0, own_key_hi = 0;

  void get_permutation(int32_t* perm, int64_t n) {
    plt_eval* e = fast_part ? fast_part.get() : this;
    const int64_t nt = e->targets();
    PLT_REQUIRE(n == nt, "get_permutation: length must be the number of target points");
    if (nt == 0) return;
    if (e->n_src == 0 || e->brute_force_branch()) {
      std::vector<int32_t> id(nt);
      for (int64_t i = 0; i < nt; ++i) id[i] = static_cast<int32_t>(i);
      PLT_CUDA(cudaMemcpyAsync(perm, id.data(), sizeof(int32_t) * nt, cudaMemcpyDefault, stream));
      PLT_CUDA(cudaStreamSynchronize(stream));
      return;
    }
    e->stream = stream;
    e->ensure_trees();
    PLT_CUDA(cudaMemcpyAsync(perm, e->target_tree().perm(), sizeof(int32_t) * nt, cudaMemcpyDefault, stream));
    PLT_CUDA(cudaStreamSynchronize(stream));
  }

  void get_shard_range(int64_t* begin, int64_t* end) {
    plt_eval* e = fast_part ? fast_part.get() : this;
    const int64_t nt = e->targets();
    if (nt == 0 || e->n_src == 0 || e->brute_force_branch()) {
      // not sharded: rank 0 computes everything (see evaluate_device)
      const bool first = e->part.on ? e->part.rank == 0 : (shard_world <= 1 || shard_rank == 0);
      *begin = 0;
      *end = first ? nt : 0;
      return;
    }
    e->stream = stream;
    e->shard_rank = shard_rank;
    e->shard_world = shard_world;
    e->arena.reset();
    e->ensure_trees();
    int leaf_lo, leaf_hi, p_lo, p_hi;
    e->shard_leaves(e->target_tree(), leaf_lo, leaf_hi, p_lo, p_hi);
    *begin = p_lo;
    *end = p_hi;
  }

  void evaluate_device(double* out) {
    const int64_t nt = targets();
    const int64_t len = kn * nt;
    if (direct_part) {
      // src/fmm/spheroidal_evaluator.hpp:24-29
      direct_part->stream = fast_part->stream = stream;
      direct_part->shard_rank = fast_part->shard_rank = shard_rank;
      direct_part->shard_world = fast_part->shard_world = shard_world;
      direct_part->timer.reset();
      fast_part->timer.reset();
      direct_part->arena.reset();
      fast_part->arena.reset();
      double* tmp = arena.take<double>(std::max<int64_t>(len, 1));
      direct_part->evaluate_device(out);
      fast_part->evaluate_device(tmp);
      if (len) PLT_LAUNCH(ctr, k_add, ceil_div(len, 256), 256, 0, stream, tmp, len, out);
      config = fast_part->config;
      return;
    }
    if (nt == 0) return;
    if (n_src == 0) {
      PLT_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * len, stream));
      config = {0, 0, kClassic};
      return;
    }
    if (kind == KIND_H && !rbf.has_hessian)
      throw Error(PLT_ERR_UNSUPPORTED, "evaluate_hessian_isotropic is not implemented for this RBF");

    const bool compact = std::isfinite(rbf.support_radius);
    // src/fmm/fmm_evaluator.hpp:226-234 / fmm_symmetric_evaluator.hpp:222-230
    const bool small = symmetric ? n_src < 1024 : n_src * nt < int64_t{1024} * 1024;
    if (force_direct || (small && !compact && force_height == 0) ||
        (compact && compact_height() <= 2 && shard_world == 1)) {
      if (shard_world > 1 || (part.on && symmetric)) {
        // brute force is not sharded: rank 0 computes it, other ranks contribute zeros
        if ((part.on ? part.rank : shard_rank) != 0) {
          PLT_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * len, stream));
          config = {0, 0, kClassic};
          return;
        }
      }
      timer.begin("direct", stream);
      brute_force(src_pos_c.get(), w_caller.get(), n_src, target_pos_c(), nt, out);
      timer.end(stream);
      config = {0, 0, kClassic};
      return;
    }

    // Source side first.  When the targets are new as well (bulk evaluation: every call) and the multipoles have to be
    // recomputed, the target tree and the plan are built on a side stream WHILE the upward pass runs on the main
    // one: both are chains of small, latency-bound kernels (0.9 + 0.45 ms against 1.0 ms on config #3; at 8 GPUs they are
    // a quarter of the step) and neither depends on the other -- the partitioned upward pass takes the cells it needs
    // from the target points themselves (ensure_partition_tables).
    timer.begin("tree", stream);
    ensure_src_tree();
    const int height = src_tree.height();
    ensure_sorted_weights();
    timer.end(stream);
    static const bool no_overlap = getenv("PLT_DEBUG_NO_OVERLAP") != nullptr;  // A/B switch
    bool upward_done = false;
    if (!no_overlap && !symmetric && !compact && height > 2 && trg_tree_stale()) {
      const plt_config c = find_best_configuration(height);
      // (partitioned: new targets need another set of source cells, ensure_trg_tree)
      if (multipole_dirty || up_order != c.order || up_d != c.d || part.on) {
        Interpolator& ip = interpolator(height, c.order, c.d);
        if (!side.s) {
          PLT_CUDA(cudaStreamCreateWithFlags(&side.s, cudaStreamNonBlocking));
          PLT_CUDA(cudaEventCreateWithFlags(&side.fork, cudaEventDisableTiming));
          PLT_CUDA(cudaEventCreateWithFlags(&side.join, cudaEventDisableTiming));
        }
        cudaStream_t side_stream = side.s;
        cudaEvent_t side_fork = side.fork, side_join = side.join;
        if (part.on) {  // (cell flags from the target points; cleared by ensure_trg_tree below, rebuilt identically)
          trg_tree.reset();
          cell_flags_valid = false;
        }
        PLT_CUDA(cudaEventRecord(side_fork, stream));  // source tree, weights and target positions are in place
        PLT_CUDA(cudaStreamWaitEvent(side_stream, side_fork, 0));
        upward(src_tree, wt_sorted.get(), ip, M, Mhat, Mblk, true, part.on);  // asynchronous, main stream
        const bool flags_ok = cell_flags_valid;
        timer.begin("tree", side_stream);
        ensure_trg_tree(side_stream);
        timer.end(side_stream);
        timer.begin("plan", side_stream);
        plan.build(src_tree, trg_tree, side_stream, ctr);
        timer.end(side_stream);
        PLT_CUDA(cudaEventRecord(side_join, side_stream));
        PLT_CUDA(cudaStreamWaitEvent(stream, side_join, 0));
        cell_flags_valid = flags_ok;  // the flags the upward pass has just used are those of this target set
        multipole_dirty = false;
        up_order = c.order;
        up_d = c.d;
        upward_done = true;
      }
    }
    if (!upward_done) {
      timer.begin("tree", stream);
      ensure_trg_tree(stream);
      timer.end(stream);
    }
    const Tree& tt = target_tree();
    if (!plan.built()) {
      timer.begin("plan", stream);
      plan.build(src_tree, tt, stream, ctr);
      timer.end(stream);
    }

    int leaf_lo, leaf_hi, p_lo, p_hi;
    shard_leaves(tt, leaf_lo, leaf_hi, p_lo, p_hi);

    double* vt = arena.take<double>(static_cast<size_t>(kn) * nt);
    PLT_CUDA(cudaMemsetAsync(vt, 0, sizeof(double) * kn * nt, stream));
    if (compact) {
      config = {height, 0, kClassic};
      timer.begin("p2p", stream);
      launch_p2p(kind, dim, rbf.k, src_tree.view(), wt_sorted.get(), tt.view(), vt, plan.view().p2p_leaves,
                 plan.n_p2p(), leaf_lo, leaf_hi, stream, ctr);
      timer.end(stream);
    } else {
      plt_config c = find_best_configuration(height);
      Interpolator& ip = interpolator(height, c.order, c.d);
      if (height > 2 && (multipole_dirty || up_order != c.order || up_d != c.d)) {
        upward(src_tree, wt_sorted.get(), ip, M, Mhat, Mblk, true, part.on);
        multipole_dirty = false;
        up_order = c.order;
        up_d = c.d;
      }
      downward(src_tree, wt_sorted.get(), Mhat, Mblk, tt, plan, ip, vt, leaf_lo, leaf_hi, true);
      config = c;
    }
    timer.begin("finish", stream);
    launch_finish_outputs(kind, dim, aniso, vt, tt.perm(), nt, p_lo, p_hi, out, stream, ctr);
    timer.end(stream);
  }
};

// =====================================================================================
// C ABI
// =====================================================================================
namespace {
thread_local std::string g_create_error;

template <class F>
int guarded(plt_eval* h, F&& f) {
  if (!h) return PLT_ERR_INVALID;
  try {
    f();
    return PLT_OK;
  } catch (const Error& e) {
    h->err = e.what();
    return e.status;
  } catch (const std::exception& e) {
    h->err = e.what();
    return PLT_ERR_INVALID;
  }
}
}  // namespace

extern "C" {

int plt_version(void) { return 200; }

int plt_tree_height(int dim, int64_t n_points) {
  if (dim < 1 || dim > 3 || n_points < 1) return 0;
  return fmm_tree_height(dim, n_points);
}

int plt_device_check(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return PLT_ERR_CUDA;
  }
  return PLT_OK;
}

int plt_eval_create(int kind, int symmetric, int dim, int rbf_id, int rbf_part, const double* params,
                    int n_params, const double* aniso, const double* bbox_min, const double* bbox_max,
                    plt_eval** out) {
  if (!out) return PLT_ERR_INVALID;
  *out = nullptr;
  auto h = std::make_unique<plt_eval>();
  try {
    h->init(kind, symmetric, dim, rbf_id, rbf_part, params, n_params, aniso, bbox_min, bbox_max);
  } catch (const Error& e) {
    g_create_error = e.what();
    return e.status;
  } catch (const std::exception& e) {
    g_create_error = e.what();
    return PLT_ERR_INVALID;
  }
  *out = h.release();
  return PLT_OK;
}

void plt_eval_destroy(plt_eval* h) { delete h; }

int plt_eval_set_source_points(plt_eval* h, const double* points, int64_t n) {
  return guarded(h, [&] {
    PLT_REQUIRE(!h->symmetric, "set_source_points is not available on a symmetric evaluator");
    h->set_points_impl(points, n, true);
  });
}

int plt_eval_set_target_points(plt_eval* h, const double* points, int64_t n) {
  return guarded(h, [&] {
    PLT_REQUIRE(!h->symmetric, "set_target_points is not available on a symmetric evaluator");
    h->set_points_impl(points, n, false);
  });
}

int plt_eval_set_points(plt_eval* h, const double* points, int64_t n) {
  return guarded(h, [&] {
    PLT_REQUIRE(h->symmetric, "set_points is only available on a symmetric evaluator");
    h->set_points_impl(points, n, true);
  });
}

int plt_eval_set_weights(plt_eval* h, const double* weights, int64_t len) {
  return guarded(h, [&] { h->set_weights(weights, len); });
}

int plt_eval_set_accuracy(plt_eval* h, double accuracy) {
  return guarded(h, [&] {
    PLT_REQUIRE(accuracy >= 0.0, "accuracy must be non-negative");
    h->set_accuracy(accuracy);
  });
}

int plt_eval_evaluate(plt_eval* h, double* out, int64_t len) {
  return guarded(h, [&] { h->evaluate(out, len); });
}

int plt_eval_evaluate_points(plt_eval* h, const double* points, int64_t n, double* out, int64_t len) {
  return guarded(h, [&] { h->evaluate_points(points, n, out, len); });
}

double plt_set_block_m2l_min_fill(double min_fill) { return blk_set_min_fill(min_fill); }

int plt_set_hadamard_tmem(int on) { return hadamard_tmem_set(on); }

int64_t plt_release_cached_memory(void) { return static_cast<int64_t>(ArenaBlockCache::get().release_all()); }

int64_t plt_cached_memory(void) { return static_cast<int64_t>(ArenaBlockCache::get().cached()); }

void plt_set_cached_memory_limit(int64_t bytes) { ArenaBlockCache::get().set_limit(bytes < 0 ? 0 : static_cast<size_t>(bytes)); }

int plt_eval_force_config(plt_eval* h, int order, int d, int tree_height_override) {
  return guarded(h, [&] {
    PLT_REQUIRE(order == 0 || (order >= 2 && order <= kMaxOrder), "order out of range");
    PLT_REQUIRE(d >= -1 && d < std::max(order, 1), "d out of range");
    auto apply = [&](plt_eval* e) {
      e->force_order = order;
      e->force_d = d;
      e->force_height = tree_height_override;
      e->best_config.clear();
    };
    apply(h);
    if (h->direct_part) {
      apply(h->fast_part.get());
    }
  });
}

int plt_eval_force_direct(plt_eval* h, int on) {
  return guarded(h, [&] {
    h->force_direct = on != 0;
    if (h->direct_part) h->direct_part->force_direct = h->fast_part->force_direct = on != 0;
  });
}

int plt_eval_get_config(plt_eval* h, plt_config* out) {
  return guarded(h, [&] {
    PLT_REQUIRE(out, "out");
    *out = h->config;
  });
}

int plt_eval_set_stream(plt_eval* h, void* cuda_stream) {
  return guarded(h, [&] { h->stream = static_cast<cudaStream_t>(cuda_stream); });
}

int plt_eval_set_target_shard(plt_eval* h, int rank, int world_size) {
  return guarded(h, [&] {
    PLT_REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, "bad shard");
    h->shard_rank = rank;
    h->shard_world = world_size;
  });
}

int plt_eval_set_partition(plt_eval* h, int rank, int world_size, int level_cut, const uint32_t* key_begin,
                           plt_allgatherv_fn allgatherv, void* ctx) {
  return guarded(h, [&] {
    auto apply = [&](plt_eval* e) {
      e->own_cells_valid = e->cell_flags_valid = false;
      e->shard_cache.valid = false;
      e->multipole_dirty = true;
      if (world_size <= 1) {
        e->part = plt_eval::Partition{};
        return;
      }
      PLT_REQUIRE(rank >= 0 && rank < world_size, "bad rank");
      PLT_REQUIRE(level_cut >= 2 && e->dim * level_cut <= 30, "the cut level must be >= 2");
      PLT_REQUIRE(key_begin != nullptr && allgatherv != nullptr, "null argument");
      for (int r = 0; r < world_size; ++r)
        PLT_REQUIRE(key_begin[r] <= key_begin[r + 1], "key ranges must be ascending");
      PLT_REQUIRE(key_begin[0] == 0, "key ranges must start at 0");
      e->part.on = true;
      e->part.rank = rank;
      e->part.world = world_size;
      e->part.cut = level_cut;
      e->part.key_begin.assign(key_begin, key_begin + world_size + 1);
      e->part.allgatherv = allgatherv;
      e->part.ctx = ctx;
    };
    PLT_REQUIRE(!std::isfinite(h->rbf.support_radius) || world_size <= 1,
                "compact-support evaluators have no far field to partition");
    apply(h);
    if (h->fast_part) apply(h->fast_part.get());
  });
}

int plt_eval_point_keys(plt_eval* h, const double* points, int64_t n, int level, uint32_t* keys) {
  return guarded(h, [&] {
    PLT_REQUIRE(n >= 0 && (n == 0 || (points && keys)), "null argument");
    PLT_REQUIRE(level >= 0 && h->dim * level <= 30, "level out of range");
    const int dim = h->dim, nside = 1 << level;
    const double inv_w = static_cast<double>(nside) / h->box.width;
    for (int64_t i = 0; i < n; ++i) {
      int c[3] = {0, 0, 0};
      for (int a = 0; a < dim; ++a) {
        double s = 0.0;  // geometry/point3d.hpp:36-40: p * A^T
        for (int b = 0; b < dim; ++b) s += points[i * dim + b] * h->aniso[a * dim + b];
        const double x = (s - (h->box.center[a] - 0.5 * h->box.width)) * inv_w;
        PLT_REQUIRE(x >= 0.0 && x <= nside, "a point lies outside the bounding box the evaluator was created with");
        c[a] = std::min(std::max(static_cast<int>(std::floor(x)), 0), nside - 1);
      }
      if (dim == 1) { int q[1] = {c[0]}; keys[i] = morton_encode<1>(q); }
      if (dim == 2) { int q[2] = {c[0], c[1]}; keys[i] = morton_encode<2>(q); }
      if (dim == 3) { int q[3] = {c[0], c[1], c[2]}; keys[i] = morton_encode<3>(q); }
    }
  });
}

int64_t plt_eval_allgather_count(plt_eval* h) {
  if (!h) return 0;
  return h->n_allgathers + (h->fast_part ? h->fast_part->n_allgathers : 0);
}

int plt_eval_get_permutation(plt_eval* h, int32_t* perm, int64_t n) {
  return guarded(h, [&] {
    PLT_REQUIRE(perm != nullptr, "null output");
    h->get_permutation(perm, n);
  });
}

int plt_eval_get_target_shard_range(plt_eval* h, int64_t* begin, int64_t* end) {
  return guarded(h, [&] {
    PLT_REQUIRE(begin && end, "null output");
    h->get_shard_range(begin, end);
  });
}

int plt_eval_gram_batched(plt_eval* h, const double* points, const int32_t* counts, int64_t n_batch, int64_t m,
                          double nugget, double* out) {
  return guarded(h, [&] {
    PLT_REQUIRE(points && counts && out, "null argument");
    PLT_REQUIRE(h->kind == KIND_K, "gram_batched needs a value-kernel (K) handle");
    PLT_REQUIRE(h->rbf_part == 0 || !h->direct_part, "gram_batched: not available on a split handle");
    PLT_REQUIRE(plt_eval::is_device_pointer(points) && plt_eval::is_device_pointer(counts) &&
                    plt_eval::is_device_pointer(out),
                "gram_batched works on device buffers");
    double a[9];
    for (int r = 0; r < h->dim; ++r)
      for (int c = 0; c < h->dim; ++c) a[r * h->dim + c] = h->aniso[r * h->dim + c];
    launch_gram_batched(h->dim, h->rbf.k, a, points, counts, n_batch, static_cast<int>(m), nugget, out, h->stream,
                        h->ctr);
  });
}

int plt_eval_gram_mixed(plt_eval* h, const double* points, const int8_t* types, int64_t n_batch, int64_t m,
                        double nugget, double* out) {
  return guarded(h, [&] {
    PLT_REQUIRE(points && types && out, "null argument");
    PLT_REQUIRE(plt_eval::is_device_pointer(points) && plt_eval::is_device_pointer(types) &&
                    plt_eval::is_device_pointer(out),
                "gram_mixed works on device buffers");
    PLT_REQUIRE(!(h->rbf_id == PLT_RBF_SPH || h->rbf_id == PLT_RBF_CUB), "Hessian of cov_spherical / cov_cubic");
    double a[9];
    for (int r = 0; r < h->dim; ++r)
      for (int c = 0; c < h->dim; ++c) a[r * h->dim + c] = h->aniso[r * h->dim + c];
    launch_gram_mixed(h->dim, h->rbf.k, a, points, reinterpret_cast<const signed char*>(types), n_batch,
                      static_cast<int>(m), nugget, out, h->stream, h->ctr);
  });
}

int plt_eval_phase_times(plt_eval* h, const char** names, double* ms, int cap) {
  if (!h) return 0;
  int n = 0;
  try {
    PLT_CUDA(cudaStreamSynchronize(h->stream));
    auto collect = [&](plt_eval* e) {
      for (auto& r : e->timer.recs) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) {
          cudaGetLastError();
          continue;
        }
        // merge phases of the same name (per-level records)
        int j = 0;
        for (; j < n; ++j)
          if (std::strcmp(names[j], r.name) == 0) break;
        if (j == n) {
          if (n == cap) continue;
          names[n] = r.name;
          ms[n] = 0.0;
          ++n;
        }
        ms[j] += t;
      }
    };
    collect(h);
    if (h->direct_part) {
      collect(h->direct_part.get());
      collect(h->fast_part.get());
    }
  } catch (const std::exception& e) {
    h->err = e.what();
  }
  return n;
}

int plt_eval_work_stats(plt_eval* h, int64_t* m2l_pairs, int64_t* m2l_target_cells, int64_t* p2p_pairs) {
  return guarded(h, [&] {
    plt_eval* e = h->fast_part ? h->fast_part.get() : h;
    unsigned long long host[3] = {0, 0, 0};
    const Tree& tt = e->target_tree();
    if (e->src_tree.built() && tt.built()) {
      DevBuf<unsigned long long> d;
      d.alloc(3, e->stream);
      d.zero(e->stream);
      LaunchCounter scratch;  // diagnostic launches are not part of the evaluation count
      launch_count_work(e->dim, e->src_tree.view(), tt.view(), d.get(), e->stream, scratch);
      PLT_CUDA(cudaMemcpyAsync(host, d.get(), sizeof(host), cudaMemcpyDeviceToHost, e->stream));
      PLT_CUDA(cudaStreamSynchronize(e->stream));
    }
    if (m2l_pairs) *m2l_pairs = static_cast<int64_t>(host[0]);
    if (m2l_target_cells) *m2l_target_cells = static_cast<int64_t>(host[1]);
    if (p2p_pairs) *p2p_pairs = static_cast<int64_t>(host[2]);
  });
}

int64_t plt_eval_launch_count(plt_eval* h) {
  if (!h) return 0;
  int64_t n = h->ctr.n;
  if (h->direct_part) n += h->direct_part->ctr.n + h->fast_part->ctr.n;
  return n;
}

const char* plt_last_error(plt_eval* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

}  // extern "C"
