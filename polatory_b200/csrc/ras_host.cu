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
#include <random>
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

unsigned hardware_threads() { return std::max(1u, std::thread::hardware_concurrency()); }

// Multi-rank selection with several threads.  On return, for every rank r of `ranks` (ascending) a[r] is the element a
// full sort would put there, everything left of it is smaller and everything right of it larger -- what successive
// std::nth_element calls on the right remainders give.  `less` must be a strict TOTAL order (ties broken by the index),
// so the sets on either side of a cut do not depend on how they were found.
// The first levels of a bisection have fewer clusters than cores; a serial selection over 10^6 items costs 40 - 75 ms
// there.  Sample sort without the sort: ~1000 sorted splitters, every thread bins its chunk (binary search) and
// scatters it bucket by bucket, then only the buckets that contain a rank are selected serially.
template <class T, class Less>
void parallel_select(T* a, size_t n, std::initializer_list<size_t> ranks, Less less, unsigned threads) {
  auto serial = [&] {
    size_t lo = 0;
    for (size_t r : ranks) {
      if (r > lo && r < n) std::nth_element(a + lo, a + r, a + n, less);
      lo = std::max(lo, std::min(r, n));
    }
  };
  threads = std::min(threads, 64u);
  if (threads <= 1 || n < (size_t{1} << 17)) return serial();
  // Thresholds that bracket every rank: from a sorted sample of m elements, the sample quantile of the rank -+ 4 standard
  // deviations of its position (0.5 sqrt(m) sample positions).  Should a rank fall outside its bracket, the selection
  // below simply runs on the (larger) class that does contain it -- the result is exact either way.
  const size_t m = 8192, margin = 4 * 46;
  std::vector<T> sample(m);
  for (size_t k = 0; k < m; ++k) sample[k] = a[(2 * k + 1) * n / (2 * m)];
  std::sort(sample.begin(), sample.end(), less);
  std::vector<size_t> cut_at;
  for (size_t r : ranks) {
    if (r == 0 || r >= n) continue;
    const size_t p = static_cast<size_t>(static_cast<double>(r) / static_cast<double>(n) * m);
    cut_at.push_back(p > margin ? p - margin : 0);
    cut_at.push_back(std::min(m - 1, p + margin));
  }
  std::sort(cut_at.begin(), cut_at.end());
  cut_at.erase(std::unique(cut_at.begin(), cut_at.end()), cut_at.end());
  std::vector<T> thr;
  for (size_t c : cut_at) thr.push_back(sample[c]);
  const size_t C = thr.size() + 1;  // classes: number of thresholds smaller than the element
  if (C == 1) return serial();
  std::unique_ptr<uint8_t[]> cls(new uint8_t[n]);
  std::vector<size_t> count(static_cast<size_t>(threads) * C, 0);
  auto chunk = [&](size_t t, size_t& lo, size_t& hi) {
    lo = n * t / threads;
    hi = n * (t + 1) / threads;
  };
  parallel_for(threads, [&](size_t t) {
    size_t lo, hi;
    chunk(t, lo, hi);
    size_t* c = count.data() + t * C;
    for (size_t k = lo; k < hi; ++k) {
      size_t b = 0;
      while (b < thr.size() && less(thr[b], a[k])) ++b;  // a handful of thresholds, the big classes first: predictable
      cls[k] = static_cast<uint8_t>(b);
      ++c[b];
    }
  });
  std::vector<size_t> start(C + 1, 0), offset(static_cast<size_t>(threads) * C);
  for (size_t b = 0; b < C; ++b) {
    size_t at = start[b];
    for (unsigned t = 0; t < threads; ++t) {
      offset[static_cast<size_t>(t) * C + b] = at;
      at += count[static_cast<size_t>(t) * C + b];
    }
    start[b + 1] = at;
  }
  std::unique_ptr<T[]> tmp(new T[n]);  // (uninitialised: its pages are first touched by the scattering threads)
  parallel_for(threads, [&](size_t t) {
    size_t lo, hi;
    chunk(t, lo, hi);
    size_t* o = offset.data() + t * C;
    for (size_t k = lo; k < hi; ++k) tmp[o[cls[k]]++] = a[k];
  });
  parallel_for(threads, [&](size_t t) {
    size_t lo, hi;
    chunk(t, lo, hi);
    std::copy(tmp.get() + lo, tmp.get() + hi, a + lo);
  });
  size_t done = 0;  // everything left of `done` is final
  for (size_t r : ranks) {
    if (r <= done || r >= n) {
      done = std::max(done, std::min(r, n));
      continue;
    }
    const size_t b = std::upper_bound(start.begin(), start.end(), r) - start.begin() - 1;  // start[b] <= r < start[b + 1]
    const size_t lo = std::max(start[b], done), hi = start[b + 1];
    std::nth_element(a + lo, a + r, a + hi, less);
    done = r;
  }
}

template <class T>
void parallel_copy(const T* src, size_t n, T* dst, unsigned threads) {
  if (threads <= 1 || n < (size_t{1} << 16)) {
    std::copy(src, src + n, dst);
    return;
  }
  parallel_for(threads, [&](size_t t) {
    const size_t lo = n * t / threads, hi = n * (t + 1) / threads;
    std::copy(src + lo, src + hi, dst + lo);
  });
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

// ---- value data: clusters and domains over COMPACT items (coordinates + index side by side) -------------------------
// The selections below touch every point of a level log2(n / leaf) times; going through idx -> pv.row(idx) is a random
// access into the point array per comparison, a compact item array is streamed.  Same order (coordinates along the
// axes, then the index), hence the same sets on either side of every cut.
struct Item {
  double x[3];
  int64_t i;
};

struct ItemLess {
  std::array<int, 3> axes;
  int dim;
  bool operator()(const Item& p, const Item& q) const {
    for (int k = 0; k < dim; ++k) {
      const int a = axes[k];
      if (p.x[a] != q.x[a]) return p.x[a] < q.x[a];
    }
    return p.i < q.i;
  }
};

template <class It>
BoxInfo box_of_items(const It* it, size_t n, int dim, unsigned threads = 1) {
  BoxInfo b;
  for (int a = 0; a < 3; ++a) {
    b.lo[a] = std::numeric_limits<double>::infinity();
    b.hi[a] = -std::numeric_limits<double>::infinity();
  }
  if (threads <= 1 || n < (size_t{1} << 16)) {
    for (size_t k = 0; k < n; ++k)
      for (int a = 0; a < dim; ++a) {
        b.lo[a] = std::min(b.lo[a], it[k].x[a]);
        b.hi[a] = std::max(b.hi[a], it[k].x[a]);
      }
  } else {
    std::vector<BoxInfo> part(threads, b);
    parallel_for(threads, [&](size_t t) {
      BoxInfo& q = part[t];
      for (size_t k = n * t / threads; k < n * (t + 1) / threads; ++k)
        for (int a = 0; a < dim; ++a) {
          q.lo[a] = std::min(q.lo[a], it[k].x[a]);
          q.hi[a] = std::max(q.hi[a], it[k].x[a]);
        }
    });
    for (const BoxInfo& q : part)
      for (int a = 0; a < dim; ++a) {
        b.lo[a] = std::min(b.lo[a], q.lo[a]);
        b.hi[a] = std::max(b.hi[a], q.hi[a]);
      }
  }
  for (int a = 0; a < 3; ++a) b.axes[a] = a;
  std::stable_sort(b.axes.begin(), b.axes.begin() + dim,
                   [&](int i, int j) { return b.hi[i] - b.lo[i] > b.hi[j] - b.lo[j]; });
  return b;
}

// A cluster is a range of ONE item array: splitting is a selection in place, no copies.
struct RangeCluster {
  size_t lo = 0, hi = 0;
  std::array<int, 3> axes{{0, 1, 2}};
  double volume = 0.0;
  int64_t centre = -1;
  size_t size() const { return hi - lo; }
};

void init_range_cluster(const Item* items, int dim, RangeCluster& c, unsigned threads = 1) {
  const Item* it = items + c.lo;
  const size_t n = c.size();
  const BoxInfo b = box_of_items(it, n, dim, threads);
  c.volume = 1.0;
  for (int a = 0; a < dim; ++a) c.volume *= b.hi[a] - b.lo[a];
  // the FIRST point nearest to the box centre (per-thread candidates are merged in position order)
  auto nearest = [&](size_t lo, size_t hi, double& best, size_t& at) {
    for (size_t k = lo; k < hi; ++k) {
      double d2 = 0.0;
      for (int a = 0; a < dim; ++a) {
        const double d = it[k].x[a] - 0.5 * (b.lo[a] + b.hi[a]);
        d2 += d * d;
      }
      if (d2 < best) {
        best = d2;
        at = k;
      }
    }
  };
  double best = std::numeric_limits<double>::infinity();
  size_t at = 0;
  if (threads <= 1 || n < (size_t{1} << 16)) {
    nearest(0, n, best, at);
  } else {
    std::vector<double> pb(threads, std::numeric_limits<double>::infinity());
    std::vector<size_t> pa(threads, 0);
    parallel_for(threads, [&](size_t t) { nearest(n * t / threads, n * (t + 1) / threads, pb[t], pa[t]); });
    for (unsigned t = 0; t < threads; ++t)
      if (pb[t] < best) {
        best = pb[t];
        at = pa[t];
      }
  }
  c.centre = n ? it[at].i : -1;
  c.axes = b.axes;
}

void split_range_cluster(Item* items, int dim, const RangeCluster& c, RangeCluster& l, RangeCluster& r,
                         unsigned threads = 1) {
  const size_t n = c.size(), mid = n > 1 ? split_position(n) : 0;
  const ItemLess less{c.axes, dim};
  Item* it = items + c.lo;
  if (mid > 0 && mid < n) parallel_select(it, n, {mid}, less, threads);
  l.lo = c.lo;
  l.hi = c.lo + mid;
  r.lo = c.lo + mid;
  r.hi = c.hi;
  // (the centre is the FIRST nearest point in the parent's sorted order; exact ties are the rule for two-point
  // clusters, so small children are put in that order before their centre is chosen)
  if (l.size() <= 8) std::sort(items + l.lo, items + l.hi, less);
  if (r.size() <= 8) std::sort(items + r.lo, items + r.hi, less);
  if (l.size()) init_range_cluster(items, dim, l, threads);
  if (r.size()) init_range_cluster(items, dim, r, threads);
}

double round_half_to_even(double d) { return std::ceil((d - 0.5) / 2.0) + std::floor((d + 0.5) / 2.0); }

}  // namespace
}  // namespace plt

using namespace plt;

struct plt_ras_domains {
  std::vector<int64_t> offsets{0};
  std::vector<int64_t> indices;
  std::vector<uint8_t> inner;
  // gradient points of the domains (Hermite data only)
  std::vector<int64_t> offsets_g;
  std::vector<int64_t> indices_g;
  std::vector<uint8_t> inner_g;
};

extern "C" {

int plt_ras_choose_coarse_points(const double* a_points, int dim, const int64_t* idcs, int64_t n_idcs,
                                 const int64_t* poly, int64_t n_poly, int64_t n_coarse, int64_t* out) {
  if (!a_points || !idcs || !out || dim < 1 || dim > 3 || n_coarse < 1) return PLT_ERR_INVALID;
  try {
    std::vector<int64_t> poly_sorted(poly, poly + n_poly);
    std::sort(poly_sorted.begin(), poly_sorted.end());
    std::vector<Item> items;
    items.reserve(n_idcs);
    for (int64_t k = 0; k < n_idcs; ++k) {
      if (std::binary_search(poly_sorted.begin(), poly_sorted.end(), idcs[k])) continue;
      Item it{{0.0, 0.0, 0.0}, idcs[k]};
      for (int a = 0; a < dim; ++a) it.x[a] = a_points[idcs[k] * dim + a];
      items.push_back(it);
    }
    if (static_cast<int64_t>(items.size()) < n_coarse) return PLT_ERR_INVALID;
    RangeCluster root;
    root.hi = items.size();
    init_range_cluster(items.data(), dim, root, hardware_threads());
    std::vector<RangeCluster> level{root};
    // Whole levels are split while the count stays below the target (the queue orders by level first).
    for (;;) {
      size_t splittable = 0;
      for (auto& c : level) splittable += c.size() > 1 ? 1 : 0;
      if (level.size() + splittable > static_cast<size_t>(n_coarse) || splittable == 0) break;
      std::vector<RangeCluster> next(level.size() * 2);
      // (the first levels have fewer clusters than cores: the cores go into the selection of each cluster)
      const unsigned inner = static_cast<unsigned>(std::max<size_t>(1, hardware_threads() / level.size()));
      parallel_for(level.size(), [&](size_t i) {
        split_range_cluster(items.data(), dim, level[i], next[2 * i], next[2 * i + 1], inner);
      });
      level.clear();
      for (auto& c : next)
        if (c.size()) level.push_back(c);
    }
    // Last, partial level: the largest boxes are split first until the target count is reached.
    std::vector<size_t> order(level.size());
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return level[x].volume > level[y].volume; });
    const size_t need = static_cast<size_t>(n_coarse) - level.size();
    std::vector<size_t> to_split;
    for (size_t k = 0; k < order.size() && to_split.size() < need; ++k)
      if (level[order[k]].size() > 1) to_split.push_back(order[k]);
    std::vector<RangeCluster> children(to_split.size() * 2);
    parallel_for(to_split.size(), [&](size_t i) {
      split_range_cluster(items.data(), dim, level[to_split[i]], children[2 * i], children[2 * i + 1]);
    });
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
    // compact items again (see Item): coordinates, index and the ownership flag side by side
    struct DItem {
      double x[3];
      int64_t i;
      uint8_t inner;
    };
    struct DBuf {  // (uninitialised storage: its pages are first touched by the threads that fill it)
      std::unique_ptr<DItem[]> p;
      size_t n = 0;
      void alloc(size_t m) {
        p.reset(new DItem[m]);
        n = m;
      }
      size_t size() const { return n; }
      DItem* data() { return p.get(); }
      const DItem* data() const { return p.get(); }
      DItem& operator[](size_t k) { return p[k]; }
      const DItem& operator[](size_t k) const { return p[k]; }
    };
    struct Dom {
      DBuf it;
    };
    std::vector<Dom> level(1), leaf_items;
    level[0].it.alloc(n_idcs);
    parallel_for(16, [&](size_t t) {
      const int64_t lo = n_idcs * static_cast<int64_t>(t) / 16, hi = n_idcs * static_cast<int64_t>(t + 1) / 16;
      for (int64_t k = lo; k < hi; ++k) {
        DItem& d = level[0].it[k];
        d.x[0] = d.x[1] = d.x[2] = 0.0;
        for (int a = 0; a < dim; ++a) d.x[a] = a_points[idcs[k] * dim + a];
        d.i = idcs[k];
        d.inner = 1;
      }
    });
    while (!level.empty()) {
      std::vector<Dom> next(level.size() * 2);
      std::vector<char> is_leaf(level.size(), 0);
      const unsigned inner_threads = static_cast<unsigned>(std::max<size_t>(1, hardware_threads() / level.size()));
      parallel_for(level.size(), [&](size_t i) {
        Dom& d = level[i];
        const int64_t n = static_cast<int64_t>(d.it.size());
        if (n <= max_leaf) {
          is_leaf[i] = 1;
          return;
        }
        const BoxInfo b = box_of_items(d.it.data(), d.it.size(), dim, inner_threads);
        // domain_divider.hpp:209-225 with unit multiplicities
        const double q = overlap_quota * static_cast<double>(max_leaf) / static_cast<double>(n);
        const int64_t n_sub = static_cast<int64_t>(round_half_to_even((1.0 + q) / 2.0 * static_cast<double>(n)));
        const int64_t left_part = n - n_sub, right_part = n_sub;
        const int64_t mid = static_cast<int64_t>(round_half_to_even(static_cast<double>(left_part + right_part) / 2.0));
        {
          // rank classes [0, left_part) [left_part, mid) [mid, right_part) [right_part, n): selections, not a sort
          const std::array<int, 3> axes = b.axes;
          auto less = [&](const DItem& p, const DItem& r) {
            for (int k = 0; k < dim; ++k) {
              const int a = axes[k];
              if (p.x[a] != r.x[a]) return p.x[a] < r.x[a];
            }
            return p.i < r.i;
          };
          parallel_select(d.it.data(), static_cast<size_t>(n),
                          {static_cast<size_t>(left_part), static_cast<size_t>(mid), static_cast<size_t>(right_part)}, less,
                          inner_threads);
        }
        Dom &l = next[2 * i], &r = next[2 * i + 1];
        l.it.alloc(right_part);
        parallel_copy(d.it.data(), static_cast<size_t>(right_part), l.it.data(), inner_threads);
        for (int64_t k = mid; k < right_part; ++k) l.it[k].inner = 0;
        r.it.alloc(n - left_part);
        parallel_copy(d.it.data() + left_part, static_cast<size_t>(n - left_part), r.it.data(), inner_threads);
        for (int64_t k = left_part; k < mid; ++k) r.it[k - left_part].inner = 0;
        d.it.p.reset();
      });
      std::vector<Dom> keep;
      for (size_t i = 0; i < level.size(); ++i) {
        if (is_leaf[i]) {
          leaf_items.push_back(std::move(level[i]));
        } else {
          keep.push_back(std::move(next[2 * i]));
          keep.push_back(std::move(next[2 * i + 1]));
        }
      }
      level.swap(keep);
    }
    struct Leaf {
      std::vector<int64_t> idx;
      std::vector<uint8_t> inner;
    };
    std::vector<Leaf> leaves(leaf_items.size());
    parallel_for(leaf_items.size(), [&](size_t i) {
      const auto& src = leaf_items[i].it;
      leaves[i].idx.resize(src.size());
      leaves[i].inner.resize(src.size());
      for (size_t k = 0; k < src.size(); ++k) {
        leaves[i].idx[k] = src[k].i;
        leaves[i].inner[k] = src[k].inner;
      }
    });
    // merge_poly_points: points sorted by index, the poly points first (inner flag carried over)
    std::vector<int64_t> poly_v(poly, poly + n_poly);
    parallel_for(leaves.size(), [&](size_t i) {
      Leaf& d = leaves[i];
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

// ---------------------------------------------------------------------------------------------
// Hermite data: value points and gradient points (multiplicity dim) mixed.  A mixed point is coded as
// idx (value point) or ~idx (gradient point).  The cut ranks are multiplicity-weighted
// (domain_divider.hpp:205-231, 66-90), so clusters are sorted (not just selected) before cutting.
// ---------------------------------------------------------------------------------------------
namespace {
struct MixedView {
  const double* p;
  const double* g;
  int dim;
  const double* row(int64_t code) const { return code >= 0 ? p + code * dim : g + (~code) * dim; }
  int mult(int64_t code) const { return code >= 0 ? 1 : dim; }
};

struct MixedBox {
  std::array<double, 3> lo, hi;
  std::array<int, 3> axes;
};

MixedBox mixed_box(const MixedView& mv, const std::vector<int64_t>& codes) {
  MixedBox b;
  for (int a = 0; a < 3; ++a) {
    b.lo[a] = std::numeric_limits<double>::infinity();
    b.hi[a] = -std::numeric_limits<double>::infinity();
  }
  for (int64_t c : codes) {
    const double* r = mv.row(c);
    for (int a = 0; a < mv.dim; ++a) {
      b.lo[a] = std::min(b.lo[a], r[a]);
      b.hi[a] = std::max(b.hi[a], r[a]);
    }
  }
  for (int a = 0; a < 3; ++a) b.axes[a] = a;
  std::stable_sort(b.axes.begin(), b.axes.begin() + mv.dim,
                   [&](int i, int j) { return b.hi[i] - b.lo[i] > b.hi[j] - b.lo[j]; });
  return b;
}

// Sorts codes (and the parallel flags, if given) lexicographically along the axes of their box.
void mixed_sort(const MixedView& mv, const std::array<int, 3>& axes, std::vector<int64_t>& codes,
                std::vector<uint8_t>* flags) {
  const size_t n = codes.size();
  std::vector<size_t> perm(n);
  std::iota(perm.begin(), perm.end(), 0);
  std::sort(perm.begin(), perm.end(), [&](size_t x, size_t y) {
    const double *p = mv.row(codes[x]), *q = mv.row(codes[y]);
    for (int k = 0; k < mv.dim; ++k) {
      const int a = axes[k];
      if (p[a] != q[a]) return p[a] < q[a];
    }
    return x < y;  // stable
  });
  std::vector<int64_t> c2(n);
  for (size_t k = 0; k < n; ++k) c2[k] = codes[perm[k]];
  codes.swap(c2);
  if (flags) {
    std::vector<uint8_t> f2(n);
    for (size_t k = 0; k < n; ++k) f2[k] = (*flags)[perm[k]];
    flags->swap(f2);
  }
}

struct MixedCluster {
  std::vector<int64_t> codes;  // sorted along the axes of the cluster's box
  double volume = 0.0;
  int64_t centre = 0;
};

void mixed_init_cluster(const MixedView& mv, MixedCluster& c) {
  const MixedBox b = mixed_box(mv, c.codes);
  c.volume = 1.0;
  for (int a = 0; a < mv.dim; ++a) c.volume *= b.hi[a] - b.lo[a];
  double best = std::numeric_limits<double>::infinity();
  c.centre = c.codes[0];
  for (int64_t code : c.codes) {  // first minimum in the incoming order, as std::min_element
    const double* r = mv.row(code);
    double d2 = 0.0;
    for (int a = 0; a < mv.dim; ++a) {
      const double d = r[a] - 0.5 * (b.lo[a] + b.hi[a]);
      d2 += d * d;
    }
    if (d2 < best) {
      best = d2;
      c.centre = code;
    }
  }
  mixed_sort(mv, b.axes, c.codes, nullptr);
}

void mixed_split_cluster(const MixedView& mv, const MixedCluster& c, MixedCluster& l, MixedCluster& r) {
  const size_t n = c.codes.size();
  std::vector<int64_t> prefix(n + 1, 0);
  for (size_t k = 0; k < n; ++k) prefix[k + 1] = prefix[k] + mv.mult(c.codes[k]);
  const int64_t total = prefix[n];
  size_t mid = 0;
  int64_t best = std::numeric_limits<int64_t>::max();
  for (size_t i = 0; i < n; ++i) {  // min_element of |2 prefix[i] - total|, ties: the even index wins
    const int64_t d = std::llabs(2 * prefix[i] - total);
    if (d < best || (d == best && i % 2 == 0)) {
      best = d;
      mid = i;
    }
  }
  l.codes.assign(c.codes.begin(), c.codes.begin() + mid);
  r.codes.assign(c.codes.begin() + mid, c.codes.end());
  if (!l.codes.empty()) mixed_init_cluster(mv, l);
  if (!r.codes.empty()) mixed_init_cluster(mv, r);
}
}  // namespace

int plt_ras_choose_coarse_points_mixed(const double* a_points, const double* a_grad_points, int dim,
                                       const int64_t* point_idcs, int64_t n_points, const int64_t* grad_idcs,
                                       int64_t n_grads, const int64_t* poly, int64_t n_poly, int64_t n_coarse_rows,
                                       int64_t* out_points, int64_t* n_out_points, int64_t* out_grads,
                                       int64_t* n_out_grads) {
  if (!a_points || !out_points || !out_grads || !n_out_points || !n_out_grads || dim < 1 || dim > 3) return PLT_ERR_INVALID;
  try {
    const MixedView mv{a_points, a_grad_points, dim};
    std::vector<int64_t> poly_sorted(poly, poly + n_poly);
    std::sort(poly_sorted.begin(), poly_sorted.end());
    MixedCluster root;
    for (int64_t k = 0; k < n_points; ++k)
      if (!std::binary_search(poly_sorted.begin(), poly_sorted.end(), point_idcs[k])) root.codes.push_back(point_idcs[k]);
    for (int64_t k = 0; k < n_grads; ++k) root.codes.push_back(~grad_idcs[k]);
    if (root.codes.empty()) return PLT_ERR_INVALID;
    mixed_init_cluster(mv, root);
    std::vector<MixedCluster> level;
    level.push_back(std::move(root));
    int64_t size = mv.mult(level[0].centre);
    std::vector<int64_t> centres_done;  // centres of clusters that stay (pop order: level, then volume)
    while (size < n_coarse_rows) {
      // children of every cluster of the level, in parallel
      std::vector<MixedCluster> next(level.size() * 2);
      parallel_for(level.size(), [&](size_t i) { mixed_split_cluster(mv, level[i], next[2 * i], next[2 * i + 1]); });
      // the queue pops the level's clusters largest box first and stops as soon as the target is reached
      std::vector<size_t> order(level.size());
      std::iota(order.begin(), order.end(), 0);
      std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return level[x].volume > level[y].volume; });
      std::vector<MixedCluster> new_level;
      std::vector<char> was_split(level.size(), 0);
      bool progress = false;
      for (size_t k : order) {
        if (size >= n_coarse_rows) break;
        was_split[k] = 1;
        size -= mv.mult(level[k].centre);
        for (MixedCluster* ch : {&next[2 * k], &next[2 * k + 1]})
          if (!ch->codes.empty()) {
            size += mv.mult(ch->centre);
            new_level.push_back(std::move(*ch));
          }
        progress = progress || level[k].codes.size() > 1;
      }
      if (size >= n_coarse_rows) {
        // final pop order: unsplit clusters of this level (by volume), then the children (by volume)
        for (size_t k : order)
          if (!was_split[k]) centres_done.push_back(level[k].centre);
        std::stable_sort(new_level.begin(), new_level.end(),
                         [](const MixedCluster& x, const MixedCluster& y) { return x.volume > y.volume; });
        for (auto& c : new_level) centres_done.push_back(c.centre);
        level.clear();
        break;
      }
      if (!progress) break;  // only singletons left: cannot reach the target
      level.swap(new_level);
    }
    for (auto& c : level) centres_done.push_back(c.centre);
    int64_t np = 0, ng = 0;
    for (int64_t k = 0; k < n_poly; ++k) out_points[np++] = poly[k];
    for (int64_t code : centres_done) {
      if (code >= 0) out_points[np++] = code;
      else out_grads[ng++] = ~code;
    }
    *n_out_points = np;
    *n_out_grads = ng;
    return PLT_OK;
  } catch (const std::exception&) {
    return PLT_ERR_INVALID;
  }
}

int plt_ras_divide_domains_mixed(const double* a_points, const double* a_grad_points, int dim, const int64_t* point_idcs,
                                 int64_t n_points, const int64_t* grad_idcs, int64_t n_grads, const int64_t* poly,
                                 int64_t n_poly, int64_t max_leaf, double overlap_quota, plt_ras_domains** out) {
  if (!a_points || !out || dim < 1 || dim > 3 || max_leaf < 2) return PLT_ERR_INVALID;
  try {
    const MixedView mv{a_points, a_grad_points, dim};
    struct Dom {
      std::vector<int64_t> codes;
      std::vector<uint8_t> inner;
    };
    std::vector<Dom> level(1), leaves;
    for (int64_t k = 0; k < n_points; ++k) level[0].codes.push_back(point_idcs[k]);
    for (int64_t k = 0; k < n_grads; ++k) level[0].codes.push_back(~grad_idcs[k]);
    level[0].inner.assign(level[0].codes.size(), 1);
    while (!level.empty()) {
      std::vector<Dom> next(level.size() * 2);
      std::vector<char> is_leaf(level.size(), 0);
      parallel_for(level.size(), [&](size_t i) {
        Dom& d = level[i];
        const int64_t n = static_cast<int64_t>(d.codes.size());
        int64_t n_mult = 0;
        for (int64_t c : d.codes) n_mult += mv.mult(c);
        if (n_mult <= max_leaf) {
          is_leaf[i] = 1;
          return;
        }
        const MixedBox b = mixed_box(mv, d.codes);
        mixed_sort(mv, b.axes, d.codes, &d.inner);
        std::vector<int64_t> prefix(n + 1, 0);
        for (int64_t k = 0; k < n; ++k) prefix[k + 1] = prefix[k] + mv.mult(d.codes[k]);
        const double q = overlap_quota * static_cast<double>(max_leaf) / static_cast<double>(n_mult);
        const int64_t n_sub = static_cast<int64_t>(round_half_to_even((1.0 + q) / 2.0 * static_cast<double>(n_mult)));
        const int64_t left_mult = n_mult - n_sub, right_mult = n_sub;
        const int64_t mid_mult = static_cast<int64_t>(round_half_to_even(static_cast<double>(left_mult + right_mult) / 2.0));
        auto ub = [&](int64_t x) {  // upper_bound(prefix, x) - 1
          return static_cast<int64_t>(std::upper_bound(prefix.begin(), prefix.end(), x) - prefix.begin()) - 1;
        };
        const int64_t left_part = ub(left_mult), right_part = ub(right_mult), mid = ub(mid_mult);
        Dom &l = next[2 * i], &r = next[2 * i + 1];
        l.codes.assign(d.codes.begin(), d.codes.begin() + right_part);
        l.inner.resize(right_part);
        for (int64_t k = 0; k < right_part; ++k) l.inner[k] = d.inner[k] && k < mid;
        r.codes.assign(d.codes.begin() + left_part, d.codes.end());
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
    std::vector<int64_t> poly_v(poly, poly + n_poly);
    auto res = std::make_unique<plt_ras_domains>();
    res->offsets_g.push_back(0);
    for (auto& d : leaves) {
      // value points sorted by index with the poly points first (merge_poly_points); gradient points in domain order
      std::vector<std::pair<int64_t, uint8_t>> pv;
      std::vector<uint8_t> front(n_poly, 0);
      for (size_t k = 0; k < d.codes.size(); ++k) {
        if (d.codes[k] < 0) {
          res->indices_g.push_back(~d.codes[k]);
          res->inner_g.push_back(d.inner[k]);
          continue;
        }
        const auto it = std::find(poly_v.begin(), poly_v.end(), d.codes[k]);
        if (it != poly_v.end()) front[it - poly_v.begin()] = d.inner[k];
        else pv.emplace_back(d.codes[k], d.inner[k]);
      }
      std::sort(pv.begin(), pv.end());
      for (int64_t k = 0; k < n_poly; ++k) {
        res->indices.push_back(poly_v[k]);
        res->inner.push_back(front[k]);
      }
      for (auto& e : pv) {
        res->indices.push_back(e.first);
        res->inner.push_back(e.second);
      }
      res->offsets.push_back(static_cast<int64_t>(res->indices.size()));
      res->offsets_g.push_back(static_cast<int64_t>(res->indices_g.size()));
    }
    *out = res.release();
    return PLT_OK;
  } catch (const std::exception&) {
    return PLT_ERR_INVALID;
  }
}

int64_t plt_ras_domains_total_grads(plt_ras_domains* h) { return h ? static_cast<int64_t>(h->indices_g.size()) : 0; }

int plt_ras_domains_get_grads(plt_ras_domains* h, int64_t* offsets, int64_t* indices, uint8_t* inner) {
  if (!h || !offsets || !indices || !inner || h->offsets_g.empty()) return PLT_ERR_INVALID;
  std::memcpy(offsets, h->offsets_g.data(), sizeof(int64_t) * h->offsets_g.size());
  std::memcpy(indices, h->indices_g.data(), sizeof(int64_t) * h->indices_g.size());
  std::memcpy(inner, h->inner_g.data(), h->inner_g.size());
  return PLT_OK;
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

// interpolation::ResidualEvaluator::set_values (include/polatory/interpolation/residual_evaluator.hpp:123-136): the
// sample of data points whose residual is measured exactly.  iota -> std::shuffle with a default-seeded
// std::mt19937 -> std::partition("value != 0" first); the caller keeps the first min(n, 1024) entries.  `block`
// doubles per point (1 for values, dim for gradient vectors, "not all zero").  Native so that the sequence is
// the standard library's own, as in the reference.
int plt_residual_sample_indices(const double* values, int64_t n, int block, int64_t* out) {
  if (n < 0 || block < 1 || (n > 0 && (!values || !out))) return PLT_ERR_INVALID;
  try {
    std::vector<int64_t> idx(static_cast<size_t>(n));
    std::iota(idx.begin(), idx.end(), int64_t{0});
    std::shuffle(idx.begin(), idx.end(), std::mt19937{});
    std::partition(idx.begin(), idx.end(), [&](int64_t i) {
      for (int c = 0; c < block; ++c)
        if (values[i * block + c] != 0.0) return true;
      return false;
    });
    std::copy(idx.begin(), idx.end(), out);
  } catch (const std::exception&) {
    return PLT_ERR_INVALID;
  }
  return PLT_OK;
}

}  // extern "C"
