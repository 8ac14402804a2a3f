// Brute-force O(N*M) near-field kernel: the reference's full_direct
// (src/fmm/full_direct.hpp:7-52), used for small problems
// (src/fmm/fmm_evaluator.hpp:226-234, fmm_symmetric_evaluator.hpp:222-230) and as the exact
// side of the accuracy search (src/fmm/fmm_accuracy_estimator.hpp:107).
//
// Mapping: one thread owns kTPT targets (registers), a CTA streams source tiles through
// shared memory (broadcast reads), the source range is split across blockIdx.y so that
// few-target/many-source shapes (the 10k x N accuracy search) still fill 148 SMs; partial
// sums are reduced in a fixed order (deterministic).
#include "direct.cuh"
#include "dispatch.cuh"

namespace plt {
namespace {

constexpr int kThreads = 128;
constexpr int kTPT = 2;
constexpr int kTile = 128;

template <int FAM, int KIND, int DIM>
__global__ void __launch_bounds__(kThreads) k_direct(DirectArgs a) {
  constexpr int KM = KindTraits<KIND, DIM>::km;
  constexpr int KN = KindTraits<KIND, DIM>::kn;
  __shared__ double s_pos[DIM][kTile];
  __shared__ double s_w[KM][kTile];
  const int tid = threadIdx.x;
  const int64_t t_base = static_cast<int64_t>(blockIdx.x) * (kThreads * kTPT);
  double tp[kTPT][DIM];
  int64_t ti[kTPT];
  double v[kTPT][KN];
#pragma unroll
  for (int u = 0; u < kTPT; ++u) {
    ti[u] = t_base + u * kThreads + tid;
    int64_t tc = ti[u] < a.nt ? ti[u] : a.nt - 1;
#pragma unroll
    for (int c = 0; c < DIM; ++c) tp[u][c] = a.tpos[c * a.nt + tc];
#pragma unroll
    for (int b = 0; b < KN; ++b) v[u][b] = 0.0;
  }
  const int64_t s0 = static_cast<int64_t>(blockIdx.y) * a.chunk;
  const int64_t s1 = s0 + a.chunk < a.ns ? s0 + a.chunk : a.ns;
  for (int64_t tile = s0; tile < s1; tile += kTile) {
    const int cnt = static_cast<int>(s1 - tile < kTile ? s1 - tile : kTile);
    __syncthreads();
    if (tid < cnt) {
#pragma unroll
      for (int c = 0; c < DIM; ++c) s_pos[c][tid] = a.spos[c * a.ns + tile + tid];
#pragma unroll
      for (int m = 0; m < KM; ++m) s_w[m][tid] = a.swt[m * a.ns + tile + tid];
    }
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < cnt; ++j) {
      double sp[DIM], w[KM];
#pragma unroll
      for (int c = 0; c < DIM; ++c) sp[c] = s_pos[c][j];
#pragma unroll
      for (int m = 0; m < KM; ++m) w[m] = s_w[m][j];
#pragma unroll
      for (int u = 0; u < kTPT; ++u) {
        if (a.symmetric && tile + j == ti[u]) continue;  // full_direct.hpp:17-19
        double d[DIM];
#pragma unroll
        for (int c = 0; c < DIM; ++c) d[c] = tp[u][c] - sp[c];
        pair_accumulate<FAM, KIND, DIM>(a.k, d, w, v[u]);
      }
    }
  }
  double* dst = a.n_chunks > 1 ? a.partial + static_cast<int64_t>(blockIdx.y) * KN * a.nt : a.out;
#pragma unroll
  for (int u = 0; u < kTPT; ++u) {
    if (ti[u] < a.nt) {
#pragma unroll
      for (int b = 0; b < KN; ++b) dst[b * a.nt + ti[u]] = v[u][b];
    }
  }
}

__global__ void k_reduce_chunks(const double* __restrict__ partial, int n_chunks, int64_t len,
                                double* __restrict__ out) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= len) return;
  double s = 0.0;
  for (int c = 0; c < n_chunks; ++c) s += partial[c * len + i];
  out[i] = s;
}

// Batched Gram matrices of the value kernel: the reference's preconditioner::mat_a
// (include/polatory/preconditioner/mat_a.hpp:10-61, value block) for B point sets at once, i.e. the
// local matrices of all RAS domains of a level in one launch.
//   out[b][i][j] = phi(A (x_bi - x_bj)) + nugget * (i == j)      i, j < count[b]
//   out[b][i][j] = (i == j)                                       otherwise (identity padding)
struct Aniso3 {
  double a[9];
};
template <int FAM, int DIM>
__global__ void __launch_bounds__(256) k_gram_batched(RbfConst k, Aniso3 A, const double* __restrict__ pts,
                                                      const int* __restrict__ counts, int m, double nugget,
                                                      double* __restrict__ out) {
  const int b = blockIdx.z;
  const int j = blockIdx.x * 16 + (threadIdx.x & 15);
  const int i = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (i >= m || j >= m) return;
  const int cnt = counts[b];
  double v;
  if (i < cnt && j < cnt) {
    const double* pi = pts + (static_cast<size_t>(b) * m + i) * DIM;
    const double* pj = pts + (static_cast<size_t>(b) * m + j) * DIM;
    double diff[DIM], d[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) diff[c] = pi[c] - pj[c];
    double r2 = 0.0;
#pragma unroll
    for (int r = 0; r < DIM; ++r) {
      d[r] = 0.0;
#pragma unroll
      for (int c = 0; c < DIM; ++c) d[r] = fma(A.a[r * DIM + c], diff[c], d[r]);
      r2 = fma(d[r], d[r], r2);
    }
    double phi = 0.0, g = 0.0, gh = 0.0;
    rbf_radial<FAM, NEED_PHI>(k, r2, phi, g, gh);
    v = phi + (i == j ? nugget : 0.0);
  } else {
    v = i == j ? 1.0 : 0.0;
  }
  out[(static_cast<size_t>(b) * m + i) * m + j] = v;
}

// Mixed value / gradient rows (Hermite data): the full mat_a of preconditioner/mat_a.hpp:10-61.
//   types[b][r] = 0: value row of point r; 1 + c: row of gradient component c; < 0: padding (identity).
//   pts[b][r][:] = coordinates of the row's point (the `dim` rows of a gradient point repeat them).
//   (value, value)      phi(x_r - x_c) + nugget (r == c)
//   (value, grad c)     -d_c phi (x_r - g_c)           = [F_iso(d) A]_c,   d = A (x_r - g_c)
//   (grad r, value)     the transpose                   = [F_iso(d) A]_r,   d = A (x_c - g_r)
//   (grad r, grad c)    -d_r d_c phi (g_r - g_c)        = [A^T H_iso(d) A]_{rc}
template <int FAM, int DIM>
__global__ void __launch_bounds__(256) k_gram_mixed(RbfConst k, Aniso3 A, const double* __restrict__ pts,
                                                    const signed char* __restrict__ types, int m, double nugget,
                                                    double* __restrict__ out) {
  const int b = blockIdx.z;
  const int c = blockIdx.x * 16 + (threadIdx.x & 15);
  const int r = blockIdx.y * 16 + (threadIdx.x >> 4);
  if (r >= m || c >= m) return;
  const int tr = types[static_cast<size_t>(b) * m + r], tc = types[static_cast<size_t>(b) * m + c];
  double v;
  if (tr < 0 || tc < 0) {
    v = r == c ? 1.0 : 0.0;
  } else {
    const double* pr = pts + (static_cast<size_t>(b) * m + r) * DIM;
    const double* pc = pts + (static_cast<size_t>(b) * m + c) * DIM;
    // value rows look at gradient columns; for (grad, value) use the transposed pair
    const bool swap = tr > 0 && tc == 0;
    double diff[DIM], d[DIM];
#pragma unroll
    for (int a = 0; a < DIM; ++a) diff[a] = swap ? pc[a] - pr[a] : pr[a] - pc[a];
#pragma unroll
    for (int i = 0; i < DIM; ++i) {
      d[i] = 0.0;
#pragma unroll
      for (int a = 0; a < DIM; ++a) d[i] = fma(A.a[i * DIM + a], diff[a], d[i]);
    }
    if (tr == 0 && tc == 0) {
      double blk[1];
      kernel_block<FAM, KIND_K, DIM>(k, d, blk);
      v = blk[0] + (r == c ? nugget : 0.0);
    } else if (tr == 0 || tc == 0) {
      const int comp = (tr == 0 ? tc : tr) - 1;
      double blk[DIM];
      kernel_block<FAM, KIND_F, DIM>(k, d, blk);
      v = 0.0;
#pragma unroll
      for (int a = 0; a < DIM; ++a) v = fma(blk[a], A.a[a * DIM + comp], v);
    } else {
      double blk[DIM * DIM];
      kernel_block<FAM, KIND_H, DIM>(k, d, blk);
      v = 0.0;
#pragma unroll
      for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int j = 0; j < DIM; ++j) v += A.a[i * DIM + (tr - 1)] * blk[i * DIM + j] * A.a[j * DIM + (tc - 1)];
    }
  }
  out[(static_cast<size_t>(b) * m + r) * m + c] = v;
}

}  // namespace

void launch_gram_mixed(int dim, const RbfConst& k, const double* aniso, const double* pts, const signed char* types,
                       int64_t n_batch, int m, double nugget, double* out, cudaStream_t stream, LaunchCounter& ctr) {
  if (n_batch == 0 || m == 0) return;
  Aniso3 A{};
  for (int i = 0; i < dim * dim; ++i) A.a[i] = aniso[i];
  PLT_REQUIRE(n_batch <= 65535, "gram_mixed: at most 65535 point sets per call");
  dim3 grid(ceil_div(m, 16), ceil_div(m, 16), static_cast<unsigned>(n_batch));
  dispatch_fkd(k.family, KIND_K, dim, [&](auto fam, auto, auto dm) {
    PLT_LAUNCH(ctr, (k_gram_mixed<fam.value, dm.value>), grid, 256, 0, stream, k, A, pts, types, m, nugget, out);
  });
}

void launch_gram_batched(int dim, const RbfConst& k, const double* aniso, const double* pts, const int* counts,
                         int64_t n_batch, int m, double nugget, double* out, cudaStream_t stream, LaunchCounter& ctr) {
  if (n_batch == 0 || m == 0) return;
  Aniso3 A{};
  for (int i = 0; i < dim * dim; ++i) A.a[i] = aniso[i];
  PLT_REQUIRE(n_batch <= 65535, "gram_batched: at most 65535 point sets per call");
  dim3 grid(ceil_div(m, 16), ceil_div(m, 16), static_cast<unsigned>(n_batch));
  dispatch_fkd(k.family, KIND_K, dim, [&](auto fam, auto, auto dm) {
    PLT_LAUNCH(ctr, (k_gram_batched<fam.value, dm.value>), grid, 256, 0, stream, k, A, pts, counts, m, nugget, out);
  });
}

int direct_plan_chunks(int64_t ns, int64_t nt) {
  const int64_t t_blocks = (nt + kThreads * kTPT - 1) / (kThreads * kTPT);
  const int64_t want = 4 * num_sm();
  int64_t chunks = (want + t_blocks - 1) / t_blocks;
  const int64_t max_chunks = std::max<int64_t>(1, ns / 1024);
  chunks = std::max<int64_t>(1, std::min(chunks, max_chunks));
  return static_cast<int>(std::min<int64_t>(chunks, 65535));
}

void launch_direct(int kind, int dim, DirectArgs a, cudaStream_t stream, LaunchCounter& ctr) {
  if (a.nt == 0) return;
  const int kn = (kind == KIND_FT || kind == KIND_H) ? dim : 1;
  if (a.ns == 0) {
    PLT_CUDA(cudaMemsetAsync(a.out, 0, sizeof(double) * kn * a.nt, stream));
    return;
  }
  PLT_REQUIRE(a.n_chunks >= 1, "n_chunks");
  a.chunk = (a.ns + a.n_chunks - 1) / a.n_chunks;
  a.chunk = (a.chunk + kTile - 1) / kTile * kTile;
  dim3 grid(ceil_div(a.nt, kThreads * kTPT), a.n_chunks);
  dispatch_fkd(a.k.family, kind, dim, [&](auto fam, auto knd, auto dm) {
    PLT_LAUNCH(ctr, (k_direct<fam.value, knd.value, dm.value>), grid, kThreads, 0, stream, a);
  });
  if (a.n_chunks > 1) {
    const int64_t len = static_cast<int64_t>(kn) * a.nt;
    PLT_LAUNCH(ctr, k_reduce_chunks, ceil_div(len, 256), 256, 0, stream, a.partial, a.n_chunks, len, a.out);
  }
}

}  // namespace plt
