// Device-resident flexible GMRES: the Krylov driver of the fit, restating
//   krylov::GmresBase  (src/krylov/gmres_base.cpp:7-85, include/polatory/krylov/gmres_base.hpp)
//   krylov::Gmres      (src/krylov/gmres.cpp:9-50: classical Gram-Schmidt Arnoldi + Givens)
//   krylov::Fgmres     (src/krylov/fgmres.cpp:8-28: x = x0 + Z y, right preconditioner only)
// as driven by interpolation::Solver::solve (include/polatory/interpolation/solver.hpp:75-142).
//
// All vectors (rhs, x0, the Krylov basis V, the preconditioned basis Z) live in HBM; the
// operator and the right preconditioner are callbacks that receive DEVICE pointers and issue
// their work on the solver's stream (the FMM matvec never leaves the device).  Per iteration
// the j+1 Arnoldi dot products are one fused pass over V (HBM-bound: 8 (j + 2) n bytes), the
// Gram-Schmidt update and the norm of the new vector are a second fused pass, and the host is
// synchronised exactly once, to fetch the new Hessenberg column for the Givens rotations
// (the (max_iter+1) x max_iter triangular factor stays on the host, as in the reference).
//
// Multi-GPU: vectors are sharded (n = local length); the dot products and norms are summed
// across ranks through the `allreduce` callback (NCCL in the harness), which is the Krylov
// reduction of SURVEY.md 8e(3).  Reductions are two-stage with a fixed order: results are
// deterministic for a fixed (n, rank count).
#include <cmath>
#include <cstring>
#include <memory>

#include "common.cuh"

namespace plt {
namespace {

constexpr int kRedBlock = 256;
constexpr int kRedGrid = 4 * kNumSM;

__device__ __forceinline__ double block_sum(double v, double* s_red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < kRedBlock / 32) t = s_red[threadIdx.x];
  if (warp == 0)
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  return t;  // valid in thread 0
}

// partial[b][i] = sum over the chunk of block b of V_i[k] * u[k],  i = 0 .. m-1
__global__ void __launch_bounds__(kRedBlock) k_multi_dot(const double* __restrict__ V, int64_t stride, int m,
                                                         const double* __restrict__ u, int64_t n,
                                                         double* __restrict__ partial) {
  __shared__ double s_red[kRedBlock / 32];
  const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = chunk * blockIdx.x, hi = min(n, lo + chunk);
  for (int i = 0; i < m; ++i) {
    const double* v = V + stride * i;
    double acc = 0.0;
    for (int64_t k = lo + threadIdx.x; k < hi; k += kRedBlock) acc = fma(v[k], u[k], acc);
    const double t = block_sum(acc, s_red);
    if (threadIdx.x == 0) partial[static_cast<size_t>(blockIdx.x) * m + i] = t;
  }
}

// out[i] = sum_b partial[b][i]  (fixed order)
__global__ void k_reduce_partials(const double* __restrict__ partial, int n_blocks, int m, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  double s = 0.0;
  for (int b = 0; b < n_blocks; ++b) s += partial[static_cast<size_t>(b) * m + i];
  out[i] = s;
}

// u -= sum_i r[i] V_i ;  partial[b] = sum over the chunk of u^2 (after the update)
__global__ void __launch_bounds__(kRedBlock) k_gs_update(const double* __restrict__ V, int64_t stride, int m,
                                                         const double* __restrict__ r, double* __restrict__ u,
                                                         int64_t n, double* __restrict__ partial) {
  __shared__ double s_red[kRedBlock / 32];
  extern __shared__ double s_r[];
  for (int i = threadIdx.x; i < m; i += kRedBlock) s_r[i] = r[i];
  __syncthreads();
  const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = chunk * blockIdx.x, hi = min(n, lo + chunk);
  double acc = 0.0;
  for (int64_t k = lo + threadIdx.x; k < hi; k += kRedBlock) {
    double x = u[k];
    // same order as the reference's loop (gmres.cpp:26-28): i ascending
    for (int i = 0; i < m; ++i) x = fma(-s_r[i], V[stride * i + k], x);
    u[k] = x;
    acc = fma(x, x, acc);
  }
  const double t = block_sum(acc, s_red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// u /= sqrt(*norm2): normalisation by a device-resident squared norm (no host round trip)
__global__ void k_scale_by_norm(double* __restrict__ u, int64_t n, const double* __restrict__ norm2) {
  const double inv = 1.0 / sqrt(*norm2);
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x)
    u[k] *= inv;
}

// y -= x  (the initial residual rhs - A x0)
__global__ void k_sub(double* __restrict__ y, const double* __restrict__ x, int64_t n) {
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x)
    y[k] -= x[k];
}

// x = x0 + sum_i y[i] Z_i
__global__ void k_combine(const double* __restrict__ Z, int64_t stride, int m, const double* __restrict__ y,
                          const double* __restrict__ x0, double* __restrict__ x, int64_t n) {
  extern __shared__ double s_r[];
  for (int i = threadIdx.x; i < m; i += blockDim.x) s_r[i] = y[i];
  __syncthreads();
  for (int64_t k = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; k < n;
       k += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    double v = x0 ? x0[k] : 0.0;
    for (int i = 0; i < m; ++i) v = fma(s_r[i], Z[stride * i + k], v);
    x[k] = v;
  }
}

}  // namespace
}  // namespace plt

using namespace plt;

struct plt_fgmres {
  int64_t n = 0;       // local length
  int64_t stride = 0;  // padded
  int max_iter = 0;
  int iter = 0;
  bool is_setup = false;
  bool x0_zero = true;
  cudaStream_t stream = nullptr;
  plt_linop_fn op = nullptr, pc = nullptr;
  void *op_ctx = nullptr, *pc_ctx = nullptr;
  plt_allreduce_fn allreduce = nullptr;
  void* ar_ctx = nullptr;
  DevBuf<double> V, Z, rhs, x0, partial, small;  // small: [max_iter + 2] reduction results / coefficients
  std::vector<double> r, c, s, g;                // host: R (column-major (max_iter+1) x max_iter), Givens, residuals
  double rhs_norm = 0.0;
  LaunchCounter ctr;
  std::string err;
  double* pinned = nullptr;

  ~plt_fgmres() {
    if (pinned) cudaFreeHost(pinned);
  }

  double& R(int i, int j) { return r[static_cast<size_t>(j) * (max_iter + 1) + i]; }
  // Dynamic shared memory of the kernels that stage m coefficients: the compiler reads them with 16-byte loads,
  // so the buffer is padded to a whole number of double pairs (compute-sanitizer flagged the 8-byte overrun).
  static size_t coeff_smem(int m) { return sizeof(double) * (static_cast<size_t>(std::max(m, 1) + 3) / 2 * 2); }
  double* v(int i) { return V.get() + stride * i; }
  double* z(int i) { return pc ? Z.get() + stride * i : v(i); }

  void init(int64_t n_, int max_iter_) {
    PLT_REQUIRE(n_ >= 0 && max_iter_ >= 1, "fgmres: bad size");
    int dev_count = 0;
    if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0) {
      cudaGetLastError();
      throw Error(PLT_ERR_CUDA, "no CUDA device");
    }
    n = n_;
    max_iter = max_iter_;
    stride = (n + 31) / 32 * 32;
    if (stride == 0) stride = 32;
    PLT_CUDA(cudaMallocHost(&pinned, sizeof(double) * (max_iter + 2)));
  }

  void apply(plt_linop_fn f, void* ctx, const double* x, double* y, const char* what) {
    const int st = f(ctx, x, y);
    if (st != 0) throw Error(PLT_ERR_INVALID, std::string("fgmres: ") + what + " callback failed");
  }

  // sums `count` device doubles across ranks (in place) and brings them to the host (synchronises)
  void finish_reduction(double* dev, int count, double* host) {
    if (allreduce) {
      const int st = allreduce(ar_ctx, dev, count);
      if (st != 0) throw Error(PLT_ERR_INVALID, "fgmres: allreduce callback failed");
    }
    PLT_CUDA(cudaMemcpyAsync(pinned, dev, sizeof(double) * count, cudaMemcpyDeviceToHost, stream));
    PLT_CUDA(cudaStreamSynchronize(stream));
    std::memcpy(host, pinned, sizeof(double) * count);
  }

  double norm2_of(const double* u_const) {
    // ||u||^2 via the update kernel with m = 0 (does not modify u)
    double* u = const_cast<double*>(u_const);
    PLT_LAUNCH(ctr, k_gs_update, kRedGrid, kRedBlock, 0, stream, V.get(), stride, 0, small.get(), u, n,
               partial.get());
    PLT_LAUNCH(ctr, k_reduce_partials, 1, 32, 0, stream, partial.get(), kRedGrid, 1, small.get());
    double h = 0.0;
    finish_reduction(small.get(), 1, &h);
    return h;
  }

  // GmresBase::GmresBase + set_initial_solution + setup (gmres_base.cpp:39-51,75-83)
  void setup(const double* rhs_in, const double* x0_in) {
    PLT_REQUIRE(op != nullptr, "fgmres: operator not set");
    V.alloc(static_cast<size_t>(stride) * (max_iter + 1), stream);
    if (pc) Z.alloc(static_cast<size_t>(stride) * max_iter, stream);
    rhs.alloc(stride, stream);
    x0.alloc(stride, stream);
    partial.alloc(static_cast<size_t>(kRedGrid) * (max_iter + 1), stream);
    small.alloc(max_iter + 2, stream);
    r.assign(static_cast<size_t>(max_iter + 1) * max_iter, 0.0);
    c.assign(max_iter, 0.0);
    s.assign(max_iter, 0.0);
    g.assign(max_iter + 1, 0.0);
    iter = 0;
    if (n) PLT_CUDA(cudaMemcpyAsync(rhs.get(), rhs_in, sizeof(double) * n, cudaMemcpyDefault, stream));
    rhs_norm = std::sqrt(norm2_of(rhs.get()));
    x0_zero = true;
    if (x0_in) {
      if (n) PLT_CUDA(cudaMemcpyAsync(x0.get(), x0_in, sizeof(double) * n, cudaMemcpyDefault, stream));
    } else {
      x0.zero(stream);
    }
    // x0_.isZero(), gmres_base.cpp:46.  With sharded vectors the decision must be the same on every rank
    // and every rank must take part in the reduction, whether or not it was given an x0 (a rank with an
    // empty shard passes none): a missing x0 counts as zeros.
    if (x0_in || allreduce) x0_zero = norm2_of(x0.get()) == 0.0;
    // r0 = rhs - A x0
    double* v0 = v(0);
    if (n) PLT_CUDA(cudaMemcpyAsync(v0, rhs.get(), sizeof(double) * n, cudaMemcpyDeviceToDevice, stream));
    if (!x0_zero) {
      double* tmp = v(1);
      apply(op, op_ctx, x0.get(), tmp, "operator");
      PLT_LAUNCH(ctr, k_sub, kRedGrid, kRedBlock, 0, stream, v0, tmp, n);
    }
    PLT_LAUNCH(ctr, k_gs_update, kRedGrid, kRedBlock, 0, stream, V.get(), stride, 0, small.get(), v0, n,
               partial.get());
    PLT_LAUNCH(ctr, k_reduce_partials, 1, 32, 0, stream, partial.get(), kRedGrid, 1, small.get());
    double n2 = 0.0;
    finish_reduction(small.get(), 1, &n2);
    g[0] = std::sqrt(n2);
    if (n2 > 0.0) PLT_LAUNCH(ctr, k_scale_by_norm, kRedGrid, kRedBlock, 0, stream, v0, n, small.get());
    is_setup = true;
  }

  // Gmres::iterate_process (gmres.cpp:9-50)
  void iterate() {
    PLT_REQUIRE(is_setup, "fgmres: setup() has not been called");
    if (iter == max_iter) return;
    const int j = iter;
    // Arnoldi: z = M^-1 v_j ; v_{j+1} = A z
    if (pc) apply(pc, pc_ctx, v(j), z(j), "preconditioner");
    double* u = v(j + 1);
    apply(op, op_ctx, z(j), u, "operator");
    // r(i, j) = <v_i, v_{j+1}>, i <= j  -- classical Gram-Schmidt, all against the un-updated vector
    const int m = j + 1;
    PLT_LAUNCH(ctr, k_multi_dot, kRedGrid, kRedBlock, 0, stream, V.get(), stride, m, u, n, partial.get());
    PLT_LAUNCH(ctr, k_reduce_partials, ceil_div(m, 128), 128, 0, stream, partial.get(), kRedGrid, m, small.get());
    if (allreduce && allreduce(ar_ctx, small.get(), m) != 0)
      throw Error(PLT_ERR_INVALID, "fgmres: allreduce callback failed");
    // v_{j+1} -= sum_i r(i, j) v_i ; r(j+1, j) = ||v_{j+1}|| ; v_{j+1} /= r(j+1, j)
    PLT_LAUNCH(ctr, k_gs_update, kRedGrid, kRedBlock, coeff_smem(m), stream, V.get(), stride, m, small.get(), u, n,
               partial.get());
    PLT_LAUNCH(ctr, k_reduce_partials, 1, 32, 0, stream, partial.get(), kRedGrid, 1, small.get() + m);
    if (allreduce && allreduce(ar_ctx, small.get() + m, 1) != 0)
      throw Error(PLT_ERR_INVALID, "fgmres: allreduce callback failed");
    PLT_LAUNCH(ctr, k_scale_by_norm, kRedGrid, kRedBlock, 0, stream, u, n, small.get() + m);
    // the one synchronisation of the iteration: the new Hessenberg column
    PLT_CUDA(cudaMemcpyAsync(pinned, small.get(), sizeof(double) * (m + 1), cudaMemcpyDeviceToHost, stream));
    PLT_CUDA(cudaStreamSynchronize(stream));
    for (int i = 0; i <= j; ++i) R(i, j) = pinned[i];
    R(j + 1, j) = std::sqrt(pinned[m]);

    // Givens rotations (gmres.cpp:32-47)
    for (int i = 0; i < j; ++i) {
      const double x = R(i, j), y = R(i + 1, j);
      R(i, j) = c[i] * x + s[i] * y;
      R(i + 1, j) = -s[i] * x + c[i] * y;
    }
    const double x = R(j, j), y = R(j + 1, j);
    const double den = std::hypot(x, y);
    c[j] = x / den;
    s[j] = y / den;
    R(j, j) = c[j] * x + s[j] * y;
    g[j + 1] = -s[j] * g[j];
    g[j] = c[j] * g[j];
    ++iter;
  }

  // Fgmres::solution_vector (fgmres.cpp:8-26)
  void solution(double* x_out) {
    PLT_REQUIRE(is_setup, "fgmres: setup() has not been called");
    std::vector<double> y(std::max(iter, 1), 0.0);
    for (int j = iter - 1; j >= 0; --j) {
      y[j] = g[j];
      for (int i = j + 1; i <= iter - 1; ++i) y[j] -= R(j, i) * y[i];
      y[j] /= R(j, j);
    }
    cudaPointerAttributes attr{};
    const bool dev_out = cudaPointerGetAttributes(&attr, x_out) == cudaSuccess &&
                         (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    DevBuf<double> tmp;
    double* dst = x_out;
    if (!dev_out) {
      tmp.alloc(stride, stream);
      dst = tmp.get();
    }
    if (iter > 0) PLT_CUDA(cudaMemcpyAsync(small.get(), y.data(), sizeof(double) * iter, cudaMemcpyHostToDevice, stream));
    PLT_LAUNCH(ctr, k_combine, kRedGrid, kRedBlock, coeff_smem(iter), stream, pc ? Z.get() : V.get(), stride, iter,
               small.get(), x0.get(), dst, n);
    if (!dev_out) {
      if (n) PLT_CUDA(cudaMemcpyAsync(x_out, dst, sizeof(double) * n, cudaMemcpyDeviceToHost, stream));
    }
    PLT_CUDA(cudaStreamSynchronize(stream));  // y (host) must outlive the copy
  }
};

namespace {
thread_local std::string g_fgmres_create_error;

template <class F>
int guarded(plt_fgmres* h, F&& f) {
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

int plt_fgmres_create(int64_t n_local, int max_iter, plt_fgmres** out) {
  if (!out) return PLT_ERR_INVALID;
  *out = nullptr;
  auto h = std::make_unique<plt_fgmres>();
  try {
    h->init(n_local, max_iter);
  } catch (const Error& e) {
    g_fgmres_create_error = e.what();
    return e.status;
  }
  *out = h.release();
  return PLT_OK;
}

void plt_fgmres_destroy(plt_fgmres* h) { delete h; }

int plt_fgmres_set_operator(plt_fgmres* h, plt_linop_fn fn, void* ctx) {
  return guarded(h, [&] {
    h->op = fn;
    h->op_ctx = ctx;
  });
}

int plt_fgmres_set_right_preconditioner(plt_fgmres* h, plt_linop_fn fn, void* ctx) {
  return guarded(h, [&] {
    PLT_REQUIRE(!h->is_setup, "fgmres: set the preconditioner before setup()");
    h->pc = fn;
    h->pc_ctx = ctx;
  });
}

int plt_fgmres_set_allreduce(plt_fgmres* h, plt_allreduce_fn fn, void* ctx) {
  return guarded(h, [&] {
    h->allreduce = fn;
    h->ar_ctx = ctx;
  });
}

int plt_fgmres_set_stream(plt_fgmres* h, void* cuda_stream) {
  return guarded(h, [&] { h->stream = static_cast<cudaStream_t>(cuda_stream); });
}

int plt_fgmres_setup(plt_fgmres* h, const double* rhs, const double* x0) {
  return guarded(h, [&] { h->setup(rhs, x0); });
}

int plt_fgmres_iterate(plt_fgmres* h) {
  return guarded(h, [&] { h->iterate(); });
}

int plt_fgmres_solution(plt_fgmres* h, double* x) {
  return guarded(h, [&] { h->solution(x); });
}

int plt_fgmres_status(plt_fgmres* h, int* iteration_count, double* absolute_residual, double* relative_residual) {
  return guarded(h, [&] {
    PLT_REQUIRE(h->is_setup, "fgmres: setup() has not been called");
    const double a = std::abs(h->g[h->iter]);
    if (iteration_count) *iteration_count = h->iter;
    if (absolute_residual) *absolute_residual = a;
    if (relative_residual) *relative_residual = a / h->rhs_norm;
  });
}

int64_t plt_fgmres_launch_count(plt_fgmres* h) { return h ? h->ctr.n : 0; }

const char* plt_fgmres_last_error(plt_fgmres* h) { return h ? h->err.c_str() : g_fgmres_create_error.c_str(); }

}  // extern "C"
