// Host-side index bookkeeping of the RAS preconditioner (no device work): the choice of the coarse
// points of a level and the recursive bisection of a level's points into overlapping domains,
// restating
//   preconditioner::DomainDivider::choose_coarse_points  include/polatory/preconditioner/domain_divider.hpp:52-123
//   preconditioner::DomainDivider::divide_domain(s)       include/polatory/preconditioner/domain_divider.hpp:171-286
//   preconditioner::Domain::merge_poly_points             include/polatory/preconditioner/domain.hpp:33-51
// for value points.  The reference walks a priority queue / a list one cluster at a time; here the work
// is level-synchronous (every cluster of a level is independent of its siblings) and spread over the
// host threads, and the reference's sort-then-cut becomes a selection at the cut ranks (same sets, O(n)).
#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstring>
#include <memory>
#include <numeric>
#include <thread>
#include <vector>

#include "common.cuh"

namespace plt {
namespace {

struct PointsView {
  const double* p;
  int dim;
  const double* row(int64_t i) const { return p + i * dim; }
};

template <class F>
void parallel_for(size_t n, F&& f) {
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const size_t nt = std::min<size_t>(hw, n);
  if (nt <= 1) {
    for (size_t i = 0; i < n; ++i) f(i);
    return;
  }
  std::atomic<size_t> next{0};
  std::vector<std::thread> pool;
  for (size_t t = 0; t < nt; ++t)
    pool.emplace_back([&] {
      for (size_t i = next++; i < n; i = next++) f(i);
    });
  for (auto& th : pool) th.join();
}

// Axes by decreasing bounding-box width (stable for equal widths), and the box itself.
struct BoxInfo {
  std::array<double, 3> lo, hi;
  std::array<int, 3> axes;
};

BoxInfo box_of(const PointsView& pv, const int64_t* idx, size_t n) {
  BoxInfo b;
  for (int a = 0; a < 3; ++a) {
    b.lo[a] = std::numeric_limits<double>::infinity();
    b.hi[a] = -std::numeric_limits<double>::infinity();
  }
  for (size_t k = 0; k < n; ++k) {
    const double* r = pv.row(idx[k]);
    for (int a = 0; a < pv.dim; ++a) {
      b.lo[a] = std::min(b.lo[a], r[a]);
      b.hi[a] = std::max(b.hi[a], r[a]);
    }
  }
  for (int a = 0; a < 3; ++a) b.axes[a] = a;
  std::stable_sort(b.axes.begin(), b.axes.begin() + pv.dim,
                   [&](int i, int j) { return b.hi[i] - b.lo[i] > b.hi[j] - b.lo[j]; });
  return b;
}

// Lexicographic order along `axes`; identical coordinates fall back to the index (deterministic).
struct AxisLess {
  const PointsView& pv;
  const std::array<int, 3>& axes;
  bool operator()(int64_t x, int64_t y) const {
    const double *p = pv.row(x), *q = pv.row(y);
    for (int k = 0; k < pv.dim; ++k) {
      const int a = axes[k];
      if (p[a] != q[a]) return p[a] < q[a];
    }
    return x < y;
  }
};

// The reference sorts a cluster / domain and then cuts it at fixed ranks; only WHICH points fall on which side
// of a cut is ever used, so a selection (nth_element) at the cut ranks replaces the sort: O(n) instead of
// O(n log n), same sets.
void select_ranks(const PointsView& pv, const std::array<int, 3>& axes, int64_t* idx, size_t n,
                  std::initializer_list<size_t> ranks) {
  const AxisLess less{pv, axes};
  size_t lo = 0;
  for (size_t r : ranks) {  // ascending ranks: each selection works on the part right of the previous cut
    if (r > lo && r < n) std::nth_element(idx + lo, idx + r, idx + n, less);
    lo = std::max(lo, std::min(r, n));
  }
}

size_t split_position(size_t size) {
  if (size % 2 == 0) return size / 2;
  const size_t a = (size - 1) / 2;  // |2 i - size| ties between a and a + 1: the even index wins
  return a % 2 == 0 ? a : a + 1;
}

struct Cluster {
  std::vector<int64_t> idx;
  std::array<int, 3> axes;  // by decreasing width of the cluster's box
  double volume;
  int64_t centre;
};

// bbox, axes and centre (the point nearest to the box centre).
void init_cluster(const PointsView& pv, Cluster& c) {
  const BoxInfo b = box_of(pv, c.idx.data(), c.idx.size());
  double best = std::numeric_limits<double>::infinity();
  c.centre = c.idx.empty() ? -1 : c.idx[0];
  c.volume = 1.0;
  for (int a = 0; a < pv.dim; ++a) c.volume *= b.hi[a] - b.lo[a];
  for (int64_t i : c.idx) {
    const double* r = pv.row(i);
    double d2 = 0.0;
    for (int a = 0; a < pv.dim; ++a) {
      const double d = r[a] - 0.5 * (b.lo[a] + b.hi[a]);
      d2 += d * d;
    }
    if (d2 < best) {
      best = d2;
      c.centre = i;
    }
  }
  c.axes = b.axes;
}

// Splits c at the reference's mid rank into (l, r).
void split_cluster(const PointsView& pv, Cluster& c, Cluster& l, Cluster& r) {
  const size_t mid = c.idx.size() > 1 ? split_position(c.idx.size()) : 0;
  select_ranks(pv, c.axes, c.idx.data(), c.idx.size(), {mid});
  l.idx.assign(c.idx.begin(), c.idx.begin() + mid);
  r.idx.assign(c.idx.begin() + mid, c.idx.end());
  if (!l.idx.empty()) init_cluster(pv, l);
  if (!r.idx.empty()) init_cluster(pv, r);
}

double round_half_to_even(double d) { return std::ceil((d - 0.5) / 2.0) + std::floor((d + 0.5) / 2.0); }

}  // namespace
}  // namespace plt

using namespace plt;

struct plt_ras_domains {
  std::vector<int64_t> offsets{0};
  std::vector<int64_t> indices;
  std::vector<uint8_t> inner;
};

extern "C" {

int plt_ras_choose_coarse_points(const double* a_points, int dim, const int64_t* idcs, int64_t n_idcs,
                                 const int64_t* poly, int64_t n_poly, int64_t n_coarse, int64_t* out) {
  if (!a_points || !idcs || !out || dim < 1 || dim > 3 || n_coarse < 1) return PLT_ERR_INVALID;
  try {
    const PointsView pv{a_points, dim};
    std::vector<int64_t> poly_sorted(poly, poly + n_poly);
    std::sort(poly_sorted.begin(), poly_sorted.end());
    Cluster root;
    root.idx.reserve(n_idcs);
    for (int64_t k = 0; k < n_idcs; ++k)
      if (!std::binary_search(poly_sorted.begin(), poly_sorted.end(), idcs[k])) root.idx.push_back(idcs[k]);
    if (static_cast<int64_t>(root.idx.size()) < n_coarse) return PLT_ERR_INVALID;
    init_cluster(pv, root);
    std::vector<Cluster> level;
    level.push_back(std::move(root));
    // Whole levels are split while the count stays below the target (the queue orders by level first).
    for (;;) {
      size_t splittable = 0;
      for (auto& c : level) splittable += c.idx.size() > 1 ? 1 : 0;
      if (level.size() + splittable > static_cast<size_t>(n_coarse) || splittable == 0) break;
      std::vector<Cluster> next(level.size() * 2);
      parallel_for(level.size(), [&](size_t i) { split_cluster(pv, level[i], next[2 * i], next[2 * i + 1]); });
      level.clear();
      for (auto& c : next)
        if (!c.idx.empty()) level.push_back(std::move(c));
    }
    // Last, partial level: the largest boxes are split first until the target count is reached.
    std::vector<size_t> order(level.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return level[x].volume > level[y].volume; });
    const size_t need = static_cast<size_t>(n_coarse) - level.size();
    std::vector<size_t> to_split;
    for (size_t k = 0; k < order.size() && to_split.size() < need; ++k)
      if (level[order[k]].idx.size() > 1) to_split.push_back(order[k]);
    std::vector<Cluster> children(to_split.size() * 2);
    parallel_for(to_split.size(),
                 [&](size_t i) { split_cluster(pv, level[to_split[i]], children[2 * i], children[2 * i + 1]); });
    std::vector<char> was_split(level.size(), 0);
    for (size_t i : to_split) was_split[i] = 1;
    int64_t w = 0;
    for (int64_t k = 0; k < n_poly; ++k) out[w++] = poly[k];
    // pop order of the reference's queue: shallower level first, larger box first
    for (size_t k : order)
      if (!was_split[k]) out[w++] = level[k].centre;
    std::vector<size_t> corder(children.size());
    std::iota(corder.begin(), corder.end(), 0);
    std::stable_sort(corder.begin(), corder.end(),
                     [&](size_t x, size_t y) { return children[x].volume > children[y].volume; });
    for (size_t k : corder) out[w++] = children[k].centre;
    return w == n_poly + n_coarse ? PLT_OK : PLT_ERR_INVALID;
  } catch (const std::exception&) {
    return PLT_ERR_INVALID;
  }
}

int plt_ras_divide_domains(const double* a_points, int dim, const int64_t* idcs, int64_t n_idcs, const int64_t* poly,
                           int64_t n_poly, int64_t max_leaf, double overlap_quota, plt_ras_domains** out) {
  if (!a_points || !idcs || !out || dim < 1 || dim > 3 || max_leaf < 2) return PLT_ERR_INVALID;
  try {
    const PointsView pv{a_points, dim};
    struct Dom {
      std::vector<int64_t> idx;
      std::vector<uint8_t> inner;
    };
    std::vector<Dom> level(1), leaves;
    level[0].idx.assign(idcs, idcs + n_idcs);
    level[0].inner.assign(n_idcs, 1);
    while (!level.empty()) {
      std::vector<Dom> next(level.size() * 2);
      std::vector<char> is_leaf(level.size(), 0);
      parallel_for(level.size(), [&](size_t i) {
        Dom& d = level[i];
        const int64_t n = static_cast<int64_t>(d.idx.size());
        if (n <= max_leaf) {
          is_leaf[i] = 1;
          return;
        }
        const BoxInfo b = box_of(pv, d.idx.data(), d.idx.size());
        // domain_divider.hpp:209-225 with unit multiplicities
        const double q = overlap_quota * static_cast<double>(max_leaf) / static_cast<double>(n);
        const int64_t n_sub = static_cast<int64_t>(round_half_to_even((1.0 + q) / 2.0 * static_cast<double>(n)));
        const int64_t left_part = n - n_sub, right_part = n_sub;
        const int64_t mid = static_cast<int64_t>(round_half_to_even(static_cast<double>(left_part + right_part) / 2.0));
        {
          // order (index, inner) pairs by rank classes [0, left_part) [left_part, mid) [mid, right_part) [right_part, n)
          std::vector<int64_t> perm(n);
          std::iota(perm.begin(), perm.end(), 0);
          const std::array<int, 3> axes = b.axes;
          auto less = [&](int64_t x, int64_t y) {
            const double *p = pv.row(d.idx[x]), *q = pv.row(d.idx[y]);
            for (int k = 0; k < dim; ++k) {
              const int a = axes[k];
              if (p[a] != q[a]) return p[a] < q[a];
            }
            return d.idx[x] < d.idx[y];
          };
          size_t lo = 0;
          for (int64_t r : {left_part, mid, right_part}) {
            if (static_cast<size_t>(r) > lo && r < n) std::nth_element(perm.begin() + lo, perm.begin() + r, perm.end(), less);
            lo = std::max<size_t>(lo, static_cast<size_t>(std::min<int64_t>(r, n)));
          }
          std::vector<int64_t> idx2(n);
          std::vector<uint8_t> in2(n);
          for (int64_t k = 0; k < n; ++k) {
            idx2[k] = d.idx[perm[k]];
            in2[k] = d.inner[perm[k]];
          }
          d.idx.swap(idx2);
          d.inner.swap(in2);
        }
        Dom &l = next[2 * i], &r = next[2 * i + 1];
        l.idx.assign(d.idx.begin(), d.idx.begin() + right_part);
        l.inner.resize(right_part);
        for (int64_t k = 0; k < right_part; ++k) l.inner[k] = d.inner[k] && k < mid;
        r.idx.assign(d.idx.begin() + left_part, d.idx.end());
        r.inner.resize(n - left_part);
        for (int64_t k = left_part; k < n; ++k) r.inner[k - left_part] = d.inner[k] && k >= mid;
      });
      std::vector<Dom> keep;
      for (size_t i = 0; i < level.size(); ++i) {
        if (is_leaf[i]) {
          leaves.push_back(std::move(level[i]));
        } else {
          keep.push_back(std::move(next[2 * i]));
          keep.push_back(std::move(next[2 * i + 1]));
        }
      }
      level.swap(keep);
    }
    // merge_poly_points: points sorted by index, the poly points first (inner flag carried over)
    std::vector<int64_t> poly_v(poly, poly + n_poly);
    parallel_for(leaves.size(), [&](size_t i) {
      Dom& d = leaves[i];
      const size_t n = d.idx.size();
      std::vector<size_t> perm(n);
      std::iota(perm.begin(), perm.end(), 0);
      std::sort(perm.begin(), perm.end(), [&](size_t x, size_t y) { return d.idx[x] < d.idx[y]; });
      std::vector<int64_t> idx2;
      std::vector<uint8_t> in2;
      idx2.reserve(n + n_poly);
      in2.reserve(n + n_poly);
      for (int64_t k = 0; k < n_poly; ++k) {
        idx2.push_back(poly_v[k]);
        in2.push_back(0);
      }
      for (size_t k = 0; k < n; ++k) {
        const int64_t id = d.idx[perm[k]];
        const auto it = std::find(poly_v.begin(), poly_v.end(), id);
        if (it != poly_v.end()) {
          in2[it - poly_v.begin()] = d.inner[perm[k]];
        } else {
          idx2.push_back(id);
          in2.push_back(d.inner[perm[k]]);
        }
      }
      d.idx.swap(idx2);
      d.inner.swap(in2);
    });
    auto res = std::make_unique<plt_ras_domains>();
    for (auto& d : leaves) {
      res->indices.insert(res->indices.end(), d.idx.begin(), d.idx.end());
      res->inner.insert(res->inner.end(), d.inner.begin(), d.inner.end());
      res->offsets.push_back(static_cast<int64_t>(res->indices.size()));
    }
    *out = res.release();
    return PLT_OK;
  } catch (const std::exception&) {
    return PLT_ERR_INVALID;
  }
}

int64_t plt_ras_domains_count(plt_ras_domains* h) { return h ? static_cast<int64_t>(h->offsets.size()) - 1 : 0; }
int64_t plt_ras_domains_total(plt_ras_domains* h) { return h ? static_cast<int64_t>(h->indices.size()) : 0; }

int plt_ras_domains_get(plt_ras_domains* h, int64_t* offsets, int64_t* indices, uint8_t* inner) {
  if (!h || !offsets || !indices || !inner) return PLT_ERR_INVALID;
  std::memcpy(offsets, h->offsets.data(), sizeof(int64_t) * h->offsets.size());
  std::memcpy(indices, h->indices.data(), sizeof(int64_t) * h->indices.size());
  std::memcpy(inner, h->inner.data(), h->inner.size());
  return PLT_OK;
}

void plt_ras_domains_destroy(plt_ras_domains* h) { delete h; }

}  // extern "C"
