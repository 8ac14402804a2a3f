// The multiplicative level sweep of the RAS preconditioner, RasPreconditioner::operator()
// (include/polatory/preconditioner/ras_preconditioner.hpp:183-246), behind the C ABI: one call applies the whole
// preconditioner to a device vector on a stream -- coarse-grid solves (coarse_grid.hpp:84-128), batched fine-level
// solves (fine_grid.hpp:103-147), level transfers through the resident evaluators (update_residuals,
// ras_preconditioner.hpp:287-321) and the orthogonalisation against the polynomials (:165-180, :267-285).
//
// The handle references the level structure the caller has set up through the other plt_ras_* / plt_chol_* entry points
// (row tables, Cholesky factors, coarse inverse, evaluator handles with their points in place); it owns only its work
// vectors.  Every row index is unique within its table, so the scatter kernels need no atomics and the sweep is
// deterministic (two-stage reductions in fixed order).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "../../include/polatory_b200.h"
#include "common.cuh"

namespace plt {
namespace {

constexpr int kT = 256;
constexpr int kDotChunks = 128;

__global__ void k_gather_fine(const double* __restrict__ res, const int64_t* __restrict__ idx,
                              const int32_t* __restrict__ cnt, int m, int64_t total, double* __restrict__ vals) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int64_t b = i / m;
  const int j = static_cast<int>(i - b * m);
  vals[i] = j < cnt[b] ? res[idx[i]] : 0.0;
}
__global__ void k_scatter_inner(const double* __restrict__ lam, const int64_t* __restrict__ glob,
                                const int64_t* __restrict__ loc, int64_t n, double* __restrict__ w) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) w[glob[i]] = lam[loc[i]];
}
__global__ void k_gather_rows(const double* __restrict__ src, const int64_t* __restrict__ rows, int64_t n,
                              double* __restrict__ dst) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) dst[i] = src[rows[i]];
}
__global__ void k_scatter_rows(const double* __restrict__ src, const int64_t* __restrict__ rows, int64_t n,
                               double* __restrict__ dst) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) dst[rows[i]] = src[i];
}
__global__ void k_sub_rows(const double* __restrict__ fit, const int64_t* __restrict__ rows, int64_t n,
                           double* __restrict__ res) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) res[rows[i]] -= fit[i];
}
// res[rows[i]] -= p_mono[rows[i]][:] . c
__global__ void k_poly_sub(const double* __restrict__ p_mono, int l, const double* __restrict__ c,
                           const int64_t* __restrict__ rows, int64_t n, double* __restrict__ res) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int64_t r = rows[i];
  double s = 0.0;
  for (int k = 0; k < l; ++k) s = fma(p_mono[r * l + k], c[k], s);
  res[r] -= s;
}
__global__ void k_add(const double* __restrict__ x, int64_t n, double* __restrict__ y) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i < n) y[i] += x[i];
}

__device__ double block_sum(double v, double* sh) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (warp == 0) {
    t = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;  // valid in warp 0
}
// partial[k][c] = sum over chunk c of p[i][k] * w[i]; grid (kDotChunks, l)
__global__ void k_dots_partial(const double* __restrict__ p, int l, const double* __restrict__ w, int64_t n,
                               double* __restrict__ partial) {
  __shared__ double sh[32];
  const int k = blockIdx.y, c = blockIdx.x;
  const int64_t per = (n + gridDim.x - 1) / gridDim.x, lo = c * per, hi = min(n, lo + per);
  double s = 0.0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) s = fma(p[i * l + k], w[i], s);
  s = block_sum(s, sh);
  if (threadIdx.x == 0) partial[k * gridDim.x + c] = s;
}
__global__ void k_dots_final(const double* __restrict__ partial, int chunks, double* __restrict__ dot) {
  __shared__ double sh[32];
  double s = 0.0;
  for (int c = threadIdx.x; c < chunks; c += blockDim.x) s += partial[blockIdx.x * chunks + c];
  s = block_sum(s, sh);
  if (threadIdx.x == 0) dot[blockIdx.x] = s;
}
// w -= p dot; res += ap dot
__global__ void k_orth_update(const double* __restrict__ p, const double* __restrict__ ap, int l,
                              const double* __restrict__ dot, int64_t n, double* __restrict__ w,
                              double* __restrict__ res) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  double a = 0.0, b = 0.0;
  for (int k = 0; k < l; ++k) {
    a = fma(p[i * l + k], dot[k], a);
    b = fma(ap[i * l + k], dot[k], b);
  }
  w[i] -= a;
  res[i] += b;
}
// t[j] = vals[l + j] + sum_k q[k][j] vals[k]
__global__ void k_coarse_pre(const double* __restrict__ vals, const double* __restrict__ q, int l, int r,
                             double* __restrict__ t) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= r) return;
  double s = vals[l + j];
  for (int k = 0; k < l; ++k) s = fma(q[static_cast<size_t>(k) * r + j], vals[k], s);
  t[j] = s;
}
// lam = [Q_top gamma; gamma]; then c = P_top^-1 (vals[:l] - A_top lam).  One CTA.
__global__ void k_coarse_post(const double* __restrict__ gamma, const double* __restrict__ q, int l, int r,
                              const double* __restrict__ vals, const double* __restrict__ a_top,
                              const double* __restrict__ p_top_inv, double* __restrict__ lam,
                              double* __restrict__ coeffs) {
  __shared__ double sh[32];
  __shared__ double rhs[16];
  const int m = l + r;
  for (int j = threadIdx.x; j < r; j += blockDim.x) lam[l + j] = gamma[j];
  for (int k = 0; k < l; ++k) {
    double s = 0.0;
    for (int j = threadIdx.x; j < r; j += blockDim.x) s = fma(q[static_cast<size_t>(k) * r + j], gamma[j], s);
    s = block_sum(s, sh);
    if (threadIdx.x == 0) lam[k] = s;
  }
  __syncthreads();
  for (int k = 0; k < l; ++k) {
    double s = 0.0;
    for (int j = threadIdx.x; j < m; j += blockDim.x) s = fma(a_top[static_cast<size_t>(k) * m + j], lam[j], s);
    s = block_sum(s, sh);
    if (threadIdx.x == 0) rhs[k] = vals[k] - s;
  }
  __syncthreads();
  if (threadIdx.x < l) {
    double s = 0.0;
    for (int k = 0; k < l; ++k) s = fma(p_top_inv[threadIdx.x * l + k], rhs[k], s);
    coeffs[threadIdx.x] = s;
  }
}

inline unsigned blocks_for(int64_t n) { return static_cast<unsigned>(std::max<int64_t>(1, (n + kT - 1) / kT)); }

}  // namespace
}  // namespace plt

using namespace plt;

struct plt_ras_sweep {
  int64_t m_rows = 0;
  int l = 0, n_levels = 0;
  struct Level {
    const int64_t* vrows = nullptr;  // value rows of the level's points
    const int64_t* grows = nullptr;  // gradient-component rows, point-major
    int64_t n_v = 0, n_g = 0;
    // fine grid (levels >= 1)
    int64_t n_dom = 0, n_inner = 0;
    int m = 0;
    const int64_t* idx = nullptr;
    const int32_t* cnt = nullptr;
    const double* factor = nullptr;
    const double* q_top = nullptr;
    const int64_t* inner_glob = nullptr;
    const int64_t* inner_loc = nullptr;
  };
  std::vector<Level> levels;
  struct Coarse {
    int m = 0;
    const int64_t* idx = nullptr;
    const double *inverse = nullptr, *q_top = nullptr, *a_top = nullptr, *p_top_inv = nullptr;
  } coarse;
  struct Transfer {
    int kind;  // 0 a, 1 f, 2 ft, 3 h
    plt_eval* ev;
  };
  std::map<std::pair<int, int>, std::vector<Transfer>> transfers;
  const double *p_mono = nullptr, *p_orth = nullptr, *a_p = nullptr;
  // work vectors
  DevBuf<double> res, w, total, vals, lam, gath, fit, small, partial;
  bool ready = false;
  std::string err;
  int64_t launches = 0;

  void check(bool ok, const char* what) {
    if (!ok) throw Error(PLT_ERR_INVALID, what);
  }

  void prepare(cudaStream_t s) {
    if (ready) return;
    check(n_levels >= 1 && static_cast<int>(levels.size()) == n_levels, "levels");
    check(coarse.idx && coarse.inverse && coarse.m > l, "the coarse grid is not set");
    check(l == 0 || (coarse.q_top && coarse.a_top && coarse.p_top_inv), "coarse polynomial tables");
    check(l <= 16, "polynomial basis too large");
    int64_t max_vals = coarse.m, max_rows = 1;
    for (int lev = 0; lev < n_levels; ++lev) {
      const Level& L = levels[lev];
      check(L.n_v == 0 || L.vrows, "level rows");
      check(L.n_g == 0 || L.grows, "level gradient rows");
      max_rows = std::max(max_rows, std::max(L.n_v, L.n_g));
      if (lev >= 1) {
        check(L.idx && L.cnt && L.factor && L.n_dom > 0 && L.m > l, "a fine level is not set");
        check(l == 0 || L.q_top, "fine level q_top");
        max_vals = std::max<int64_t>(max_vals, L.n_dom * L.m);
      }
    }
    if (n_levels > 1 && l > 0) check(p_mono && p_orth && a_p, "polynomial tables");
    res.alloc(m_rows, s);
    w.alloc(m_rows + l, s);
    total.alloc(m_rows + l, s);
    vals.alloc(max_vals, s);
    lam.alloc(max_vals, s);
    gath.alloc(max_rows, s);
    fit.alloc(max_rows, s);
    small.alloc(64 + 2 * static_cast<size_t>(coarse.m), s);
    partial.alloc(static_cast<size_t>(kDotChunks) * std::max(l, 1), s);
    ready = true;
  }

  // CoarseGrid::solve + set_solution_to (coarse_grid.hpp:84-128): w = 0 except the coarse rows and the coefficients
  void solve_coarse(cudaStream_t s) {
    const int m = coarse.m, r = m - l;
    PLT_CUDA(cudaMemsetAsync(w.get(), 0, sizeof(double) * (m_rows + l), s));
    k_gather_rows<<<blocks_for(m), kT, 0, s>>>(res.get(), coarse.idx, m, vals.get());
    double* t = small.get() + 64;          // [r]
    double* gamma = small.get() + 64 + m;  // [r]
    if (l > 0) {
      k_coarse_pre<<<blocks_for(r), kT, 0, s>>>(vals.get(), coarse.q_top, l, r, t);
      if (plt_gemv(coarse.inverse, r, r, t, gamma, s) != PLT_OK) throw Error(PLT_ERR_CUDA, "coarse gemv");
      k_coarse_post<<<1, kT, 0, s>>>(gamma, coarse.q_top, l, r, vals.get(), coarse.a_top, coarse.p_top_inv, lam.get(),
                                     w.get() + m_rows);
      launches += 3;
    } else {
      if (plt_gemv(coarse.inverse, m, m, vals.get(), lam.get(), s) != PLT_OK) throw Error(PLT_ERR_CUDA, "coarse gemv");
      launches += 1;
    }
    k_scatter_rows<<<blocks_for(m), kT, 0, s>>>(lam.get(), coarse.idx, m, w.get());
    launches += 2;
  }

  // FineGrid::solve + set_solution_to for every domain of the level (fine_grid.hpp:103-147)
  void solve_fine(int lev, cudaStream_t s) {
    const Level& L = levels[lev];
    const int64_t tot = L.n_dom * L.m;
    PLT_CUDA(cudaMemsetAsync(w.get(), 0, sizeof(double) * (m_rows + l), s));
    k_gather_fine<<<blocks_for(tot), kT, 0, s>>>(res.get(), L.idx, L.cnt, L.m, tot, vals.get());
    if (plt_chol_solve_batched(L.factor, L.n_dom, L.m - l, L.q_top, l, vals.get(), lam.get(), s) != PLT_OK)
      throw Error(PLT_ERR_CUDA, "batched fine-level solve");
    k_scatter_inner<<<blocks_for(L.n_inner), kT, 0, s>>>(lam.get(), L.inner_glob, L.inner_loc, L.n_inner, w.get());
    launches += 3;
  }
  void solve(int lev, cudaStream_t s) {
    if (lev == 0) solve_coarse(s);
    else solve_fine(lev, s);
  }

  // update_residuals (ras_preconditioner.hpp:287-321): res[target level rows] -= A(target, source) w[source level rows]
  void update_residuals(int src, int trg, cudaStream_t s) {
    const Level &S = levels[src], &T = levels[trg];
    auto it = transfers.find({src, trg});
    if (it != transfers.end()) {
      for (const Transfer& t : it->second) {
        const bool src_grad = t.kind == 1 || t.kind == 3, trg_grad = t.kind == 2 || t.kind == 3;
        const int64_t ns = src_grad ? S.n_g : S.n_v, nt = trg_grad ? T.n_g : T.n_v;
        if (ns == 0 || nt == 0) continue;
        k_gather_rows<<<blocks_for(ns), kT, 0, s>>>(w.get(), src_grad ? S.grows : S.vrows, ns, gath.get());
        if (plt_eval_set_stream(t.ev, s) != PLT_OK || plt_eval_set_weights(t.ev, gath.get(), ns) != PLT_OK ||
            plt_eval_evaluate(t.ev, fit.get(), nt) != PLT_OK)
          throw Error(PLT_ERR_CUDA, plt_last_error(t.ev));
        k_sub_rows<<<blocks_for(nt), kT, 0, s>>>(fit.get(), trg_grad ? T.grows : T.vrows, nt, res.get());
        launches += 2;
      }
    }
    if (l > 0) {
      if (T.n_v) k_poly_sub<<<blocks_for(T.n_v), kT, 0, s>>>(p_mono, l, w.get() + m_rows, T.vrows, T.n_v, res.get());
      if (T.n_g) k_poly_sub<<<blocks_for(T.n_g), kT, 0, s>>>(p_mono, l, w.get() + m_rows, T.grows, T.n_g, res.get());
      launches += 2;
    }
  }
  void add_total(cudaStream_t s) {
    k_add<<<blocks_for(m_rows + l), kT, 0, s>>>(w.get(), m_rows + l, total.get());
    ++launches;
  }
  // orthogonalize (ras_preconditioner.hpp:267-285): total.head -= P (P^T total.head); res += A P (P^T total.head)
  void orthogonalize(cudaStream_t s) {
    if (l == 0) return;
    double* dot = small.get();
    k_dots_partial<<<dim3(kDotChunks, l), kT, 0, s>>>(p_orth, l, total.get(), m_rows, partial.get());
    k_dots_final<<<l, kT, 0, s>>>(partial.get(), kDotChunks, dot);
    k_orth_update<<<blocks_for(m_rows), kT, 0, s>>>(p_orth, a_p, l, dot, m_rows, total.get(), res.get());
    launches += 3;
  }

  void apply(const double* v, double* out, cudaStream_t s) {
    prepare(s);
    const int n = n_levels;
    PLT_CUDA(cudaMemcpyAsync(res.get(), v, sizeof(double) * m_rows, cudaMemcpyDeviceToDevice, s));
    if (n == 1) {
      solve(0, s);
      PLT_CUDA(cudaMemcpyAsync(out, w.get(), sizeof(double) * (m_rows + l), cudaMemcpyDeviceToDevice, s));
      return;
    }
    PLT_CUDA(cudaMemsetAsync(total.get(), 0, sizeof(double) * (m_rows + l), s));
    solve(0, s);
    update_residuals(0, n - 1, s);
    add_total(s);
    for (int lev = 1; lev < n - 1; ++lev) {
      solve(lev, s);
      update_residuals(lev, n - 1, s);
      add_total(s);
      orthogonalize(s);
      solve(0, s);
      update_residuals(0, n - 1, s);
      add_total(s);
    }
    for (int lev = n - 1; lev >= 1; --lev) {
      solve(lev, s);
      update_residuals(lev, lev - 1, s);
      add_total(s);
      orthogonalize(s);
      solve(0, s);
      if (lev > 1) update_residuals(0, lev - 1, s);
      add_total(s);
    }
    PLT_CUDA(cudaMemcpyAsync(out, total.get(), sizeof(double) * (m_rows + l), cudaMemcpyDeviceToDevice, s));
    PLT_CUDA(cudaGetLastError());
  }
};

namespace {
template <class F>
int sweep_guarded(plt_ras_sweep* h, F&& f) {
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

int plt_ras_sweep_create(int64_t m_rows, int l, int n_levels, plt_ras_sweep** out) {
  if (!out || m_rows < 1 || l < 0 || n_levels < 1) return PLT_ERR_INVALID;
  auto h = new plt_ras_sweep();
  h->m_rows = m_rows;
  h->l = l;
  h->n_levels = n_levels;
  h->levels.resize(n_levels);
  *out = h;
  return PLT_OK;
}

void plt_ras_sweep_destroy(plt_ras_sweep* h) { delete h; }

int plt_ras_sweep_set_level_rows(plt_ras_sweep* h, int level, const int64_t* value_rows, int64_t n_value,
                                 const int64_t* grad_rows, int64_t n_grad) {
  return sweep_guarded(h, [&] {
    h->check(level >= 0 && level < h->n_levels && n_value >= 0 && n_grad >= 0, "level");
    auto& L = h->levels[level];
    L.vrows = value_rows;
    L.n_v = n_value;
    L.grows = grad_rows;
    L.n_g = n_grad;
    h->ready = false;
  });
}

int plt_ras_sweep_set_fine(plt_ras_sweep* h, int level, int64_t n_domains, int m, const int64_t* idx, const int32_t* cnt,
                           const double* factor, const double* q_top, const int64_t* inner_glob,
                           const int64_t* inner_loc, int64_t n_inner) {
  return sweep_guarded(h, [&] {
    h->check(level >= 1 && level < h->n_levels && n_domains > 0 && m > h->l && n_inner >= 0, "fine level");
    auto& L = h->levels[level];
    L.n_dom = n_domains;
    L.m = m;
    L.idx = idx;
    L.cnt = cnt;
    L.factor = factor;
    L.q_top = q_top;
    L.inner_glob = inner_glob;
    L.inner_loc = inner_loc;
    L.n_inner = n_inner;
    h->ready = false;
  });
}

int plt_ras_sweep_set_coarse(plt_ras_sweep* h, int m, const int64_t* idx, const double* inverse, const double* q_top,
                             const double* a_top, const double* p_top_inv) {
  return sweep_guarded(h, [&] {
    h->check(m > h->l && idx && inverse, "coarse grid");
    h->coarse.m = m;
    h->coarse.idx = idx;
    h->coarse.inverse = inverse;
    h->coarse.q_top = q_top;
    h->coarse.a_top = a_top;
    h->coarse.p_top_inv = p_top_inv;
    h->ready = false;
  });
}

int plt_ras_sweep_add_transfer(plt_ras_sweep* h, int src_level, int trg_level, int kind, plt_eval* ev) {
  return sweep_guarded(h, [&] {
    h->check(src_level >= 0 && src_level < h->n_levels && trg_level >= 0 && trg_level < h->n_levels, "levels");
    h->check(kind >= 0 && kind <= 3 && ev, "transfer");
    h->transfers[{src_level, trg_level}].push_back({kind, ev});
  });
}

int plt_ras_sweep_set_poly(plt_ras_sweep* h, const double* p_mono, const double* p_orth, const double* a_p) {
  return sweep_guarded(h, [&] {
    h->p_mono = p_mono;
    h->p_orth = p_orth;
    h->a_p = a_p;
    h->ready = false;
  });
}

int plt_ras_sweep_apply(plt_ras_sweep* h, const double* v, double* out, void* stream) {
  return sweep_guarded(h, [&] {
    h->check(v && out, "vectors");
    h->apply(v, out, static_cast<cudaStream_t>(stream));
  });
}

int64_t plt_ras_sweep_launch_count(plt_ras_sweep* h) { return h ? h->launches : 0; }

const char* plt_ras_sweep_last_error(plt_ras_sweep* h) { return h ? h->err.c_str() : "null handle"; }

}  // extern "C"
