// Device construction of the interaction plan (plan.cuh).
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "plan.cuh"

namespace plt {
namespace {

template <int DIM>
struct Stencil {
  static constexpr int NN = DIM == 1 ? 3 : (DIM == 2 ? 9 : 27);
  static constexpr int NC = 1 << DIM;
  static constexpr int CENTER = (NN - 1) / 2;
};

// flags[g - cell_off[1]] = 1 for every target cell g of levels 1 .. leaf-1 that has a source
// cell among its 3^dim - 1 non-central same-level neighbours (its children then have a
// non-empty M2L list; the central neighbour only contributes near-field pairs).
template <int DIM>
__global__ void k_plan_mark(TreeView src, TreeView trg, int* __restrict__ flags) {
  const int leaf = trg.height - 1;
  const int g = trg.cell_off[1] + blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= trg.cell_off[leaf]) return;
  int l = 1;
  while (l + 1 < leaf && g >= trg.cell_off[l + 1]) ++l;
  int c[DIM];
  morton_decode<DIM>(trg.keys[g], c);
  const int nside = 1 << l;
  const int* sd = src.dense + src.dense_off[l];
  int found = 0;
  for (int e = 0; e < Stencil<DIM>::NN && !found; ++e) {
    if (e == Stencil<DIM>::CENTER) continue;
    int q[DIM], r = e;
    bool ok = true;
#pragma unroll
    for (int a = DIM - 1; a >= 0; --a) {
      q[a] = c[a] + (r % 3) - 1;
      r /= 3;
      ok = ok && q[a] >= 0 && q[a] < nside;
    }
    if (ok && sd[morton_encode<DIM>(q)] >= 0) found = 1;
  }
  flags[g - trg.cell_off[1]] = found;
}

template <int DIM>
__global__ void k_plan_p2p_mark(TreeView src, TreeView trg, int* __restrict__ flags) {
  const int leaf = trg.height - 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= trg.n_cells[leaf]) return;
  int c[DIM];
  morton_decode<DIM>(trg.keys[trg.cell_off[leaf] + i], c);
  const int nside = 1 << leaf;
  const int* sd = src.dense + src.dense_off[leaf];
  int found = 0;
  for (int e = 0; e < Stencil<DIM>::NN && !found; ++e) {
    int q[DIM], r = e;
    bool ok = true;
#pragma unroll
    for (int a = DIM - 1; a >= 0; --a) {
      q[a] = c[a] + (r % 3) - 1;
      r /= 3;
      ok = ok && q[a] >= 0 && q[a] < nside;
    }
    if (ok && sd[morton_encode<DIM>(q)] >= 0) found = 1;
  }
  flags[i] = found;
}

__device__ __forceinline__ int level_of_parent(const TreeView& trg, int g) {
  const int leaf = trg.height - 1;
  int l = 1;
  while (l + 1 < leaf && g >= trg.cell_off[l + 1]) ++l;
  return l;
}

// Sibling groups (plan.cuh): flags[i] = 1 if the i-th active parent starts a run of parents with a common parent.
__global__ void k_plan_grp_flags(TreeView trg, const int* __restrict__ raw, const int* __restrict__ counts, int n_par,
                                 int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_par) return;
  int f = 0;
  if (i < counts[0]) {
    f = 1;
    if (i > 0) {
      const int g = raw[i] + trg.cell_off[1], gp = raw[i - 1] + trg.cell_off[1];
      f = level_of_parent(trg, g) != level_of_parent(trg, gp) || (trg.keys[g] >> trg.dim) != (trg.keys[gp] >> trg.dim);
    }
  }
  flags[i] = f;
}

// counts: [0] n_active, [1] n_p2p, [2 + l] level_begin[l] for l = 0 .. height, [27] n_groups,
// [28 + l] grp_level_begin[l] (grp_first == nullptr: no groups).
constexpr int kCountGroups = 27, kCountsSize = 28 + 25;
__global__ void k_plan_bounds(TreeView trg, const int* __restrict__ raw, const int* __restrict__ grp_first,
                              int* __restrict__ counts) {
  const int l = threadIdx.x;
  if (l > trg.height) return;
  const int n = counts[0];
  int v;
  if (l < 2) {
    v = 0;
  } else if (l >= trg.height) {
    v = n;
  } else {
    // first active entry whose parent level is >= l - 1
    const int key = trg.cell_off[l - 1] - trg.cell_off[1];
    int lo = 0, hi = n;
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (raw[mid] < key) lo = mid + 1; else hi = mid;
    }
    v = lo;
  }
  counts[2 + l] = v;
  int gv = 0;
  if (grp_first) {  // first group whose first slot is >= level_begin[l] (a group never spans two levels)
    int lo = 0, hi = counts[kCountGroups];
    while (lo < hi) {
      int mid = (lo + hi) >> 1;
      if (grp_first[mid] < v) lo = mid + 1; else hi = mid;
    }
    gv = lo;
  }
  counts[kCountGroups + 1 + l] = gv;
}

// grp_slot / grp_src (plan.cuh), 3-D: one thread per (group, position of the 4^3 source block).
__global__ void k_plan_grp_fill(TreeView src, TreeView trg, const int* __restrict__ raw, const int* __restrict__ grp_first,
                                int n_groups, int n_active, int* __restrict__ grp_slot, int* __restrict__ grp_src) {
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int grp = static_cast<int>(t >> 6), e = static_cast<int>(t & 63);
  if (grp >= n_groups) return;
  const int head = grp_first[grp];
  const int next = grp + 1 < n_groups ? grp_first[grp + 1] : n_active;
  const int g = raw[head] + trg.cell_off[1];
  const int pl = level_of_parent(trg, g);
  int gc[3];
  morton_decode<3>(trg.keys[g] >> 3, gc);
  const int nside = 1 << pl;
  int q[3] = {2 * gc[0] + (e >> 4) - 1, 2 * gc[1] + ((e >> 2) & 3) - 1, 2 * gc[2] + (e & 3) - 1};
  int id = -1;
  if (q[0] >= 0 && q[0] < nside && q[1] >= 0 && q[1] < nside && q[2] >= 0 && q[2] < nside) {
    const int ci = src.dense[src.dense_off[pl] + morton_encode<3>(q)];
    if (ci >= 0) id = src.cell_off[pl] + ci - src.cell_off[1];
  }
  grp_src[static_cast<size_t>(grp) * 64 + e] = id;
  if (e < 8) {
    int slot = -1;
    for (int j = head; j < next; ++j)
      if ((trg.keys[raw[j] + trg.cell_off[1]] & 7u) == static_cast<unsigned>(e)) slot = j;
    grp_slot[static_cast<size_t>(grp) * 8 + e] = slot;
  }
}

struct LevelBegin {
  int v[25];
};

template <int DIM>
__global__ void k_plan_fill(TreeView src, TreeView trg, const int* __restrict__ raw, LevelBegin lb, int n_active,
                            int* __restrict__ active, int* __restrict__ src_ids,
                            unsigned char* __restrict__ trg_mask, int* __restrict__ leaf_slot) {
  constexpr int NN = Stencil<DIM>::NN, NC = Stencil<DIM>::NC;
  const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int slot = static_cast<int>(t / (NN * NC));
  const int e = static_cast<int>(t % (NN * NC));
  if (slot >= n_active) return;
  int l = 2;  // level of the children
  while (l + 1 < trg.height && slot >= lb.v[l + 1]) ++l;
  const int pl = l - 1;
  const int g = raw[slot] + trg.cell_off[1];
  const int pidx = g - trg.cell_off[pl];
  const uint32_t pkey = trg.keys[g];
  int pc[DIM];
  morton_decode<DIM>(pkey, pc);
  const int nside_p = 1 << pl;
  const int nb = e / NC, ch = e % NC;
  int q[DIM], r = nb;
  bool ok = true;
#pragma unroll
  for (int d = DIM - 1; d >= 0; --d) {
    q[d] = pc[d] + (r % 3) - 1;
    r /= 3;
    ok = ok && q[d] >= 0 && q[d] < nside_p;
  }
  int id = -1;
  if (ok) {
    const uint32_t ck = (morton_encode<DIM>(q) << DIM) | ch;
    const int ci = src.dense[src.dense_off[l] + ck];
    if (ci >= 0) id = src.cell_off[l] + ci - src.cell_off[2];
  }
  src_ids[static_cast<size_t>(slot) * (NN * NC) + e] = id;
  if (e == 0) {
    active[slot] = pidx;
    unsigned m = 0;
    const int* td = trg.dense + trg.dense_off[l];
    for (int c = 0; c < NC; ++c)
      if (td[(pkey << DIM) | c] >= 0) m |= 1u << c;
    trg_mask[slot] = static_cast<unsigned char>(m);
    if (l == trg.height - 1) leaf_slot[pidx] = slot - lb.v[l];
  }
}

// leaf_meta (plan.cuh): one thread per (parent of the leaf level, child).
template <int DIM>
__global__ void k_plan_leaf_meta(TreeView trg, const int* __restrict__ leaf_slot, int* __restrict__ meta) {
  constexpr int NC = Stencil<DIM>::NC, W = 2 + 3 * NC;
  const int leaf = trg.height - 1, pl = leaf - 1;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int pidx = t / NC, ch = t % NC;
  if (pidx >= trg.n_cells[pl]) return;
  const uint32_t pkey = trg.keys[trg.cell_off[pl] + pidx];
  int* m = meta + static_cast<size_t>(pidx) * W;
  if (ch == 0) {
    m[0] = static_cast<int>(pkey);
    m[1] = leaf_slot[pidx];
  }
  const int cidx = trg.dense[trg.dense_off[leaf] + ((pkey << DIM) | ch)];
  int first = 0, cnt = 0;
  if (cidx >= 0) {
    first = trg.leaf_start[cidx];
    cnt = trg.leaf_start[cidx + 1] - first;
  }
  m[2 + 3 * ch] = cidx;
  m[3 + 3 * ch] = first;
  m[4 + 3 * ch] = cnt;
}

}  // namespace

void Plan::build(const Tree& src, const Tree& trg, cudaStream_t stream, LaunchCounter& ctr) {
  const TreeView sv = src.view(), tv = trg.view();
  PLT_REQUIRE(sv.height == tv.height && sv.dim == tv.dim, "source and target trees must have the same shape");
  const int dim = tv.dim, height = tv.height, leaf = height - 1;
  const int nn = dim == 1 ? 3 : (dim == 2 ? 9 : 27), nc = 1 << dim;
  view_ = PlanView{};
  counts_.alloc(kCountsSize, stream);
  counts_.zero(stream);

  // ---- phase 1: flags + compaction (M2L parents of all levels at once; P2P leaves) ----
  const int n_par = height > 2 ? tv.cell_off[leaf] - tv.cell_off[1] : 0;
  size_t tmp_bytes = 0;
  thrust::counting_iterator<int> iota(0);
  if (n_par > 0) {
    flags_.alloc(n_par, stream);
    active_.alloc(2 * static_cast<size_t>(n_par), stream);  // [raw | level-local ids]
    if (dim == 1) PLT_LAUNCH(ctr, k_plan_mark<1>, ceil_div(n_par, 256), 256, 0, stream, sv, tv, flags_.get());
    if (dim == 2) PLT_LAUNCH(ctr, k_plan_mark<2>, ceil_div(n_par, 256), 256, 0, stream, sv, tv, flags_.get());
    if (dim == 3) PLT_LAUNCH(ctr, k_plan_mark<3>, ceil_div(n_par, 256), 256, 0, stream, sv, tv, flags_.get());
    PLT_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, iota, flags_.get(), active_.get(), counts_.get(), n_par,
                                        stream));
    if (tmp_bytes > tmp_.size()) tmp_.alloc(tmp_bytes, stream);
    PLT_CUDA(cub::DeviceSelect::Flagged(tmp_.get(), tmp_bytes, iota, flags_.get(), active_.get(), counts_.get(), n_par,
                                        stream));
    ctr.n += 1;
    if (dim == 3) {  // sibling groups of the active parents
      grp_flags_.alloc(n_par, stream);
      grp_first_.alloc(n_par, stream);
      PLT_LAUNCH(ctr, k_plan_grp_flags, ceil_div(n_par, 256), 256, 0, stream, tv, active_.get(), counts_.get(), n_par,
                 grp_flags_.get());
      PLT_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, iota, grp_flags_.get(), grp_first_.get(),
                                          counts_.get() + kCountGroups, n_par, stream));
      if (tmp_bytes > tmp_.size()) tmp_.alloc(tmp_bytes, stream);
      PLT_CUDA(cub::DeviceSelect::Flagged(tmp_.get(), tmp_bytes, iota, grp_flags_.get(), grp_first_.get(),
                                          counts_.get() + kCountGroups, n_par, stream));
      ctr.n += 1;
    }
  }
  const bool grouped = n_par > 0 && dim == 3;
  const int n_leaf = tv.n_cells[leaf];
  p2p_flags_.alloc(n_leaf, stream);
  p2p_leaves_.alloc(n_leaf, stream);
  if (dim == 1) PLT_LAUNCH(ctr, k_plan_p2p_mark<1>, ceil_div(n_leaf, 256), 256, 0, stream, sv, tv, p2p_flags_.get());
  if (dim == 2) PLT_LAUNCH(ctr, k_plan_p2p_mark<2>, ceil_div(n_leaf, 256), 256, 0, stream, sv, tv, p2p_flags_.get());
  if (dim == 3) PLT_LAUNCH(ctr, k_plan_p2p_mark<3>, ceil_div(n_leaf, 256), 256, 0, stream, sv, tv, p2p_flags_.get());
  PLT_CUDA(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, iota, p2p_flags_.get(), p2p_leaves_.get(), counts_.get() + 1,
                                      n_leaf, stream));
  if (tmp_bytes > tmp_.size()) tmp_.alloc(tmp_bytes, stream);
  PLT_CUDA(cub::DeviceSelect::Flagged(tmp_.get(), tmp_bytes, iota, p2p_flags_.get(), p2p_leaves_.get(),
                                      counts_.get() + 1, n_leaf, stream));
  ctr.n += 1;
  PLT_LAUNCH(ctr, k_plan_bounds, 1, 32, 0, stream, tv, active_.get(), grouped ? grp_first_.get() : nullptr,
             counts_.get());

  int host[kCountsSize];
  PLT_CUDA(cudaMemcpyAsync(host, counts_.get(), sizeof(host), cudaMemcpyDeviceToHost, stream));
  PLT_CUDA(cudaStreamSynchronize(stream));  // the one synchronisation of the plan
  const int n_active = host[0];
  LevelBegin lb{};
  for (int l = 0; l <= height; ++l) {
    lb.v[l] = host[2 + l];
    view_.level_begin[l] = host[2 + l];
  }
  view_.n_active = n_active;
  view_.n_groups = grouped ? host[kCountGroups] : 0;
  for (int l = 0; l <= height; ++l) view_.grp_level_begin[l] = grouped ? host[kCountGroups + 1 + l] : 0;
  view_.n_p2p = host[1];
  view_.p2p_leaves = p2p_leaves_.get();

  // ---- phase 2: resolve source ids, child masks and the leaf slot map ----
  if (height > 2) {
    leaf_slot_.alloc(tv.n_cells[leaf - 1], stream);
    leaf_slot_.fill_byte(0xFF, stream);
    src_ids_.alloc(static_cast<size_t>(std::max(n_active, 1)) * nn * nc, stream);
    trg_mask_.alloc(std::max(n_active, 1), stream);
    int* active_out = active_.get() + n_par;
    if (n_active > 0) {
      const int64_t threads = static_cast<int64_t>(n_active) * nn * nc;
      if (dim == 1) PLT_LAUNCH(ctr, k_plan_fill<1>, ceil_div(threads, 256), 256, 0, stream, sv, tv, active_.get(), lb, n_active, active_out, src_ids_.get(), trg_mask_.get(), leaf_slot_.get());
      if (dim == 2) PLT_LAUNCH(ctr, k_plan_fill<2>, ceil_div(threads, 256), 256, 0, stream, sv, tv, active_.get(), lb, n_active, active_out, src_ids_.get(), trg_mask_.get(), leaf_slot_.get());
      if (dim == 3) PLT_LAUNCH(ctr, k_plan_fill<3>, ceil_div(threads, 256), 256, 0, stream, sv, tv, active_.get(), lb, n_active, active_out, src_ids_.get(), trg_mask_.get(), leaf_slot_.get());
    }
    if (view_.n_groups > 0) {
      grp_slot_.alloc(static_cast<size_t>(view_.n_groups) * 8, stream);
      grp_src_.alloc(static_cast<size_t>(view_.n_groups) * 64, stream);
      PLT_LAUNCH(ctr, k_plan_grp_fill, ceil_div(static_cast<int64_t>(view_.n_groups) * 64, 256), 256, 0, stream, sv, tv,
                 active_.get(), grp_first_.get(), view_.n_groups, n_active, grp_slot_.get(), grp_src_.get());
      view_.grp_first = grp_first_.get();
      view_.grp_slot = grp_slot_.get();
      view_.grp_src = grp_src_.get();
    }
    view_.active = active_out;
    view_.src_ids = src_ids_.get();
    view_.trg_mask = trg_mask_.get();
    view_.leaf_slot = leaf_slot_.get();
    const int n_lp = tv.n_cells[leaf - 1];
    leaf_meta_.alloc(static_cast<size_t>(std::max(n_lp, 1)) * (2 + 3 * nc), stream);
    if (n_lp > 0) {
      const int threads = n_lp * nc;
      if (dim == 1) PLT_LAUNCH(ctr, k_plan_leaf_meta<1>, ceil_div(threads, 256), 256, 0, stream, tv, leaf_slot_.get(), leaf_meta_.get());
      if (dim == 2) PLT_LAUNCH(ctr, k_plan_leaf_meta<2>, ceil_div(threads, 256), 256, 0, stream, tv, leaf_slot_.get(), leaf_meta_.get());
      if (dim == 3) PLT_LAUNCH(ctr, k_plan_leaf_meta<3>, ceil_div(threads, 256), 256, 0, stream, tv, leaf_slot_.get(), leaf_meta_.get());
    }
    view_.leaf_meta = leaf_meta_.get();
  }
  built_ = true;
}

}  // namespace plt
