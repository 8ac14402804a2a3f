// Dense local problems of the RAS preconditioner on the device, hand-written (no cuSOLVER / cuBLAS):
//
//   preconditioner::FineGrid::setup / solve     include/polatory/preconditioner/fine_grid.hpp:59-142
//   preconditioner::CoarseGrid::setup / solve   include/polatory/preconditioner/coarse_grid.hpp:40-131
//
// The reference factorises Q^T A Q of every domain with Eigen's LDLT, spills the factor to a temp file
// (binary_cache.hpp) and re-reads it for each solve.  Here every domain of a level is one matrix of a batch:
//   k_reduce_q          Q^T A Q = A_rr + Q_top^T (A_tt Q_top + A_tr) + A_rt Q_top       (fine_grid.hpp:71-81)
//   k_chol_batched      blocked right-looking Cholesky, one CTA per matrix, factor kept in HBM
//   k_chol_solve        Q^T d, the two triangular solves, lambda = Q gamma, one CTA per domain (fine_grid.hpp:112-133)
// Matrices are row-major [batch][n][n]; the factor overwrites the lower triangle (L L^T, L lower).  Domains smaller
// than the batch's n are padded with identity rows / zero right-hand sides by the caller (ras.py).
#include <cmath>
#include <cstdint>

#include "common.cuh"

namespace plt {
namespace {

constexpr int kNB = 32;        // panel width
constexpr int kTS = 64;        // trailing-update tile
constexpr int kCholThreads = 256;

// red[b][i][j] = a_rr[i][j] + sum_s q[s][i] (sum_t a_tt[s][t] q[t][j] + a_tr[s][j]) + sum_s a_rt[i][s] q[s][j]
// a: [B][m][m], q: [B][l][r], r = m - l, red: [B][r][r].  One thread per element; l <= 10.
__global__ void k_reduce_q(const double* __restrict__ a, const double* __restrict__ q, int m, int l,
                           double* __restrict__ red) {
  const int r = m - l;
  const int b = blockIdx.z;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r || j >= r) return;
  const double* ab = a + static_cast<size_t>(b) * m * m;
  const double* qb = q + static_cast<size_t>(b) * l * r;
  double acc = ab[static_cast<size_t>(l + i) * m + (l + j)];
  for (int s = 0; s < l; ++s) {
    double t = ab[static_cast<size_t>(s) * m + (l + j)];                       // a_tr[s][j]
    for (int u = 0; u < l; ++u) t = fma(ab[static_cast<size_t>(s) * m + u], qb[static_cast<size_t>(u) * r + j], t);
    acc = fma(qb[static_cast<size_t>(s) * r + i], t, acc);
    acc = fma(ab[static_cast<size_t>(l + i) * m + s], qb[static_cast<size_t>(s) * r + j], acc);  // a_rt[i][s] q[s][j]
  }
  red[(static_cast<size_t>(b) * r + i) * r + j] = acc;
}

// Blocked right-looking Cholesky of one n x n matrix per CTA, in place.  Output format (internal to the solve
// kernel below): off-diagonal blocks hold L in the lower triangle AND L^T mirrored in the upper triangle (so that a
// column of L is a contiguous row segment); the kNB x kNB diagonal blocks hold L11^-1 (symmetric fill).
//   for each panel of kNB columns: (1) factorise the diagonal block in shared memory, (2) triangular-solve the
//   rows below it (one thread per row), (3) rank-kNB update of the trailing lower triangle in kTS x kTS tiles
//   (both panel blocks staged in shared memory, 4 x 4 outputs per thread).
__global__ void __launch_bounds__(kCholThreads) k_chol_batched(double* __restrict__ a_all, int n, int* __restrict__ info) {
  __shared__ double D[kNB][kNB + 1];
  __shared__ __align__(16) double PI[kNB][kTS + 4];  // panel block of the tile's rows, k-major
  double (*Dinv)[kNB + 1] = reinterpret_cast<double (*)[kNB + 1]>(&PI[0][0]);  // PI is idle during the block step
  __shared__ __align__(16) double PJ[kNB][kTS + 4];  // panel block of the tile's columns
  __shared__ int s_bad;
  double* a = a_all + static_cast<size_t>(blockIdx.x) * n * n;
  const int tid = threadIdx.x;
  if (tid == 0) s_bad = 0;
  for (int k0 = 0; k0 < n; k0 += kNB) {
    const int nb = min(kNB, n - k0);
    // ---- (1) diagonal block ----
    __syncthreads();
    for (int e = tid; e < nb * nb; e += kCholThreads) {
      const int i = e / nb, j = e - i * nb;
      D[i][j] = j <= i ? a[static_cast<size_t>(k0 + i) * n + (k0 + j)] : 0.0;
    }
    __syncthreads();
    for (int j = 0; j < nb; ++j) {
      if (tid == 0) {
        const double p = D[j][j];
        if (!(p > 0.0) && s_bad == 0) s_bad = k0 + j + 1;  // not positive definite (or NaN)
        D[j][j] = sqrt(p);
      }
      __syncthreads();
      const double inv = 1.0 / D[j][j];
      if (tid > j && tid < nb) D[tid][j] *= inv;
      __syncthreads();
      // trailing part of the block: D[i][c] -= D[i][j] D[c][j], j < c <= i < nb
      const int w = nb - j - 1;
      for (int e = tid; e < w * w; e += kCholThreads) {
        const int i = j + 1 + e / w, c = j + 1 + e % w;
        if (c <= i) D[i][c] = fma(-D[i][j], D[c][j], D[i][c]);
      }
      __syncthreads();
    }
    // The solves only ever apply L11^-1 (a 32 x 32 triangular matrix-vector product, no dependent chain), so the
    // diagonal block is stored INVERTED: column j of L11^-1 by forward substitution, one thread per column.
    if (tid < nb) {
      const int j = tid;
      for (int i = j; i < nb; ++i) {
        double sum = i == j ? 1.0 : 0.0;
        for (int t = j; t < i; ++t) sum = fma(-D[i][t], Dinv[t][j], sum);
        Dinv[i][j] = sum / D[i][i];
      }
    }
    __syncthreads();
    for (int e = tid; e < nb * nb; e += kCholThreads) {
      const int i = e / nb, j = e - i * nb;
      a[static_cast<size_t>(k0 + i) * n + (k0 + j)] = j <= i ? Dinv[i][j] : Dinv[j][i];  // symmetric fill
    }
    const int r0 = k0 + nb;   // first row below the panel's diagonal block
    if (r0 >= n) break;
    // ---- (2) panel: L21[i][:] = A21[i][:] L11^-T, kTS rows at a time staged k-major in shared memory (coalesced
    // row segments in, conflict-free column accesses), one thread per row for the substitution ----
    for (int i0 = r0; i0 < n; i0 += 2 * kTS) {
      __syncthreads();
      for (int e = tid; e < 2 * kTS * kNB; e += kCholThreads) {
        const int rr = e / kNB, k = e - rr * kNB;
        const double v = (i0 + rr < n && k < nb) ? a[static_cast<size_t>(i0 + rr) * n + (k0 + k)] : 0.0;
        if (rr < kTS) PI[k][rr] = v; else PJ[k][rr - kTS] = v;
      }
      __syncthreads();
      if (tid < 2 * kTS) {
        double (*Pt)[kTS + 4] = tid < kTS ? PI : PJ;
        const int rr = tid < kTS ? tid : tid - kTS;
        for (int c = 0; c < nb; ++c) {
          double s = Pt[c][rr];
          for (int t = 0; t < c; ++t) s = fma(-Pt[t][rr], D[c][t], s);
          Pt[c][rr] = s / D[c][c];
        }
      }
      __syncthreads();
      for (int e = tid; e < 2 * kTS * kNB; e += kCholThreads) {
        const int rr = e / kNB, k = e - rr * kNB;
        if (i0 + rr < n && k < nb)
          a[static_cast<size_t>(i0 + rr) * n + (k0 + k)] = rr < kTS ? PI[k][rr] : PJ[k][rr - kTS];
      }
      // mirror into the upper triangle (row k0 + k, columns i0 ...: contiguous): the solves read L by columns
      for (int e = tid; e < 2 * kTS * kNB; e += kCholThreads) {
        const int k = e / (2 * kTS), rr = e - k * (2 * kTS);
        if (i0 + rr < n && k < nb)
          a[static_cast<size_t>(k0 + k) * n + (i0 + rr)] = rr < kTS ? PI[k][rr] : PJ[k][rr - kTS];
      }
    }
    __syncthreads();
    // ---- (3) trailing update of the lower triangle: A[I][J] -= P[I] P[J]^T ----
    const int nt = (n - r0 + kTS - 1) / kTS;
    const int ty = tid / 16, tx = tid % 16;
    for (int ti = 0; ti < nt; ++ti) {
      for (int tj = 0; tj <= ti; ++tj) {
        const int i0 = r0 + ti * kTS, j0 = r0 + tj * kTS;
        __syncthreads();
        for (int e = tid; e < kTS * kNB; e += kCholThreads) {
          const int rr = e / kNB, k = e - rr * kNB;   // row-contiguous global reads (kNB doubles per row)
          PI[k][rr] = (i0 + rr < n && k < nb) ? a[static_cast<size_t>(i0 + rr) * n + (k0 + k)] : 0.0;
          PJ[k][rr] = (j0 + rr < n && k < nb) ? a[static_cast<size_t>(j0 + rr) * n + (k0 + k)] : 0.0;
        }
        __syncthreads();
        double c[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int v = 0; v < 4; ++v) c[u][v] = 0.0;
#pragma unroll 8
        for (int k = 0; k < kNB; ++k) {
          // 32 contiguous bytes per thread (two LDS.128): conflict-free across the 16 tx lanes, broadcast across ty
          const double2* pi2 = reinterpret_cast<const double2*>(&PI[k][ty * 4]);
          const double2* pj2 = reinterpret_cast<const double2*>(&PJ[k][tx * 4]);
          const double2 i01 = pi2[0], i23 = pi2[1], j01 = pj2[0], j23 = pj2[1];
          const double pi[4] = {i01.x, i01.y, i23.x, i23.y}, pj[4] = {j01.x, j01.y, j23.x, j23.y};
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) c[u][v] = fma(pi[u], pj[v], c[u][v]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int gi = i0 + ty * 4 + u;
          if (gi >= n) continue;
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const int gj = j0 + tx * 4 + v;
            if (gj <= gi && gj < n) a[static_cast<size_t>(gi) * n + gj] -= c[u][v];
          }
        }
      }
    }
  }
  __syncthreads();
  if (tid == 0 && info) info[blockIdx.x] = s_bad;
}

// One domain per CTA:  qtd = vals[l:] + Q_top^T vals[:l];  L y = qtd;  L^T gamma = y;  lambda = [Q_top gamma; gamma].
// F: [B][n][n] factor in the format written by k_chol_batched; q: [B][l][n] or null (l == 0); vals, lam: [B][l + n].
// Blocked substitution: the diagonal block is applied as y_blk = L11^-1 x_blk (one lane per row, no dependent chain),
// the remaining rows are updated from the MIRRORED factor (thread per row, coalesced across the CTA).
constexpr int kSolveThreads = 256;
__global__ void __launch_bounds__(kSolveThreads) k_chol_solve(const double* __restrict__ F_all, size_t f_stride, int n,
                                                              const double* __restrict__ q_all, int l,
                                                              const double* __restrict__ vals_all,
                                                              double* __restrict__ lam_all) {
  extern __shared__ double sm[];
  double* x = sm;                      // [n] right-hand side / solution
  double* Dg = x + ((n + 31) & ~31);   // [kNB][kNB + 1] inverse of the diagonal block
  double* yb = Dg + kNB * (kNB + 1);   // [kNB] solved block
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = kSolveThreads / 32;
  const double* F = F_all + static_cast<size_t>(b) * f_stride;  // f_stride == 0: one factor, many right-hand sides
  const double* q = l ? q_all + static_cast<size_t>(b) * l * n : nullptr;
  const double* vals = vals_all + static_cast<size_t>(b) * (l + n);
  double* lam = lam_all + static_cast<size_t>(b) * (l + n);
  for (int i = tid; i < n; i += kSolveThreads) {
    double v = vals[l + i];
    for (int s = 0; s < l; ++s) v = fma(q[static_cast<size_t>(s) * n + i], vals[s], v);
    x[i] = v;
  }
  // ---- forward: L y = x ----
  for (int k0 = 0; k0 < n; k0 += kNB) {
    const int nb = min(kNB, n - k0);
    __syncthreads();
    for (int e = tid; e < nb * nb; e += kSolveThreads) {
      const int i = e / nb, j = e - i * nb;
      Dg[i * (kNB + 1) + j] = F[static_cast<size_t>(k0 + i) * n + (k0 + j)];
    }
    __syncthreads();
    if (warp == 0) {
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int c = 0; c < kNB; c += 2) {  // entries above the diagonal of L11^-1 are multiplied by 0
        a0 = fma(c <= lane && c < nb ? Dg[lane * (kNB + 1) + c] : 0.0, c < nb ? x[k0 + c] : 0.0, a0);
        a1 = fma(c + 1 <= lane && c + 1 < nb ? Dg[lane * (kNB + 1) + c + 1] : 0.0, c + 1 < nb ? x[k0 + c + 1] : 0.0, a1);
      }
      yb[lane] = a0 + a1;
    }
    __syncthreads();
    if (tid < nb) x[k0 + tid] = yb[tid];
    // rows below: x[i] -= sum_c L[i][k0 + c] y[c],  L[i][k0 + c] = F[k0 + c][i] (mirror): coalesced over i
    // (a full panel below a partial one cannot occur: the partial panel is the last)
    for (int i = k0 + kNB + tid; i < n; i += kSolveThreads) {
      const double* col = F + static_cast<size_t>(k0) * n + i;
      double v[kNB];
#pragma unroll
      for (int c = 0; c < kNB; ++c) v[c] = col[static_cast<size_t>(c) * n];  // kNB independent coalesced loads in flight
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int c = 0; c < kNB; c += 2) {
        s0 = fma(v[c], yb[c], s0);
        s1 = fma(v[c + 1], yb[c + 1], s1);
      }
      x[i] -= s0 + s1;
    }
  }
  // ---- backward: L^T gamma = y ----
  const int last = ((n - 1) / kNB) * kNB;
  for (int k0 = last; k0 >= 0; k0 -= kNB) {
    const int nb = min(kNB, n - k0);
    __syncthreads();
    // x[k0 + j] -= sum_{i >= k0 + nb} L[i][k0 + j] gamma[i] = sum_i F[k0 + j][i] gamma[i]: warp per j, lanes over i
    for (int j = warp; j < nb; j += NW) {
      const double* rowj = F + static_cast<size_t>(k0 + j) * n;
      double acc = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
      int i = k0 + nb + lane;
      for (; i + 224 < n; i += 256) {  // 8 independent coalesced loads in flight per lane
        const double v0 = rowj[i], v1 = rowj[i + 32], v2 = rowj[i + 64], v3 = rowj[i + 96];
        const double v4 = rowj[i + 128], v5 = rowj[i + 160], v6 = rowj[i + 192], v7 = rowj[i + 224];
        acc = fma(v0, x[i], acc); acc1 = fma(v1, x[i + 32], acc1); acc2 = fma(v2, x[i + 64], acc2);
        acc3 = fma(v3, x[i + 96], acc3); acc = fma(v4, x[i + 128], acc); acc1 = fma(v5, x[i + 160], acc1);
        acc2 = fma(v6, x[i + 192], acc2); acc3 = fma(v7, x[i + 224], acc3);
      }
      for (; i < n; i += 32) acc = fma(rowj[i], x[i], acc);
      acc += acc1 + acc2 + acc3;
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) yb[j] = x[k0 + j] - acc;
    }
    for (int e = tid; e < nb * nb; e += kSolveThreads) {
      const int i = e / nb, j = e - i * nb;
      Dg[i * (kNB + 1) + j] = F[static_cast<size_t>(k0 + i) * n + (k0 + j)];
    }
    __syncthreads();
    // gamma_blk = L11^-T x_blk: lane r sums over c >= r of L11^-1[c][r] x[c]
    if (warp == 0) {
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int c = 0; c < kNB; c += 2) {
        a0 = fma(c >= lane && c < nb ? Dg[c * (kNB + 1) + lane] : 0.0, c < nb ? yb[c] : 0.0, a0);
        a1 = fma(c + 1 >= lane && c + 1 < nb ? Dg[(c + 1) * (kNB + 1) + lane] : 0.0, c + 1 < nb ? yb[c + 1] : 0.0, a1);
      }
      if (lane < nb) x[k0 + lane] = a0 + a1;
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += kSolveThreads) lam[l + i] = x[i];
  // lambda_top = Q_top gamma
  for (int s = warp; s < l; s += NW) {
    double acc = 0.0;
    for (int i = lane; i < n; i += 32) acc = fma(q[static_cast<size_t>(s) * n + i], x[i], acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) lam[s] = acc;
  }
}

// y = A x, A row-major [rows][cols]: one warp per row (the coarse grid's explicit inverse, coarse_grid.hpp:84-128).
__global__ void k_gemv(const double* __restrict__ A, int rows, int cols, const double* __restrict__ x,
                       double* __restrict__ y) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const double* a = A + static_cast<size_t>(row) * cols;
  double acc0 = 0.0, acc1 = 0.0;
  int j = lane;
  for (; j + 32 < cols; j += 64) {
    acc0 = fma(a[j], x[j], acc0);
    acc1 = fma(a[j + 32], x[j + 32], acc1);
  }
  if (j < cols) acc0 = fma(a[j], x[j], acc0);
  acc0 += acc1;
  for (int o = 16; o > 0; o >>= 1) acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
  if (lane == 0) y[row] = acc0;
}

bool is_dev(const void* p) {
  cudaPointerAttributes attr{};
  const bool dev = cudaPointerGetAttributes(&attr, p) == cudaSuccess &&
                   (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
  cudaGetLastError();
  return dev;
}

}  // namespace
}  // namespace plt

using namespace plt;

extern "C" {

int plt_ras_reduce_q(const double* a, const double* q_top, int64_t n_batch, int m, int l, double* red, void* stream) {
  if (!a || !red || n_batch < 0 || l < 0 || m <= l || (l > 0 && !q_top)) return PLT_ERR_INVALID;
  if (!is_dev(a) || !is_dev(red) || (l > 0 && !is_dev(q_top))) return PLT_ERR_INVALID;
  if (n_batch == 0) return PLT_OK;
  const int r = m - l;
  auto s = static_cast<cudaStream_t>(stream);
  const dim3 block(32, 8);
  for (int64_t b0 = 0; b0 < n_batch; b0 += 65535) {
    const int nb = static_cast<int>(std::min<int64_t>(65535, n_batch - b0));
    const dim3 grid((r + 31) / 32, (r + 7) / 8, nb);
    k_reduce_q<<<grid, block, 0, s>>>(a + static_cast<size_t>(b0) * m * m, q_top ? q_top + static_cast<size_t>(b0) * l * r : nullptr,
                                      m, l, red + static_cast<size_t>(b0) * r * r);
  }
  return cudaGetLastError() == cudaSuccess ? PLT_OK : PLT_ERR_CUDA;
}

int plt_chol_batched(double* a, int64_t n_batch, int n, int* info, void* stream) {
  if (!a || n_batch < 0 || n < 1) return PLT_ERR_INVALID;
  if (!is_dev(a) || (info && !is_dev(info))) return PLT_ERR_INVALID;
  if (n_batch == 0) return PLT_OK;
  k_chol_batched<<<static_cast<unsigned>(n_batch), kCholThreads, 0, static_cast<cudaStream_t>(stream)>>>(a, n, info);
  return cudaGetLastError() == cudaSuccess ? PLT_OK : PLT_ERR_CUDA;
}

int plt_gemv(const double* a, int rows, int cols, const double* x, double* y, void* stream) {
  if (!a || !x || !y || rows < 1 || cols < 1) return PLT_ERR_INVALID;
  if (!is_dev(a) || !is_dev(x) || !is_dev(y)) return PLT_ERR_INVALID;
  k_gemv<<<(rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, rows, cols, x, y);
  return cudaGetLastError() == cudaSuccess ? PLT_OK : PLT_ERR_CUDA;
}

int plt_chol_solve_shared(const double* factor, int n, int64_t n_rhs, const double* rhs, double* out, void* stream) {
  if (!factor || !rhs || !out || n < 1 || n_rhs < 0) return PLT_ERR_INVALID;
  if (!is_dev(factor) || !is_dev(rhs) || !is_dev(out)) return PLT_ERR_INVALID;
  if (n_rhs == 0) return PLT_OK;
  const size_t smem = sizeof(double) * (((n + 31) & ~31) + kNB * (kNB + 1) + kNB);
  if (smem > 40 * 1024 &&
      cudaFuncSetAttribute((const void*)k_chol_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) !=
          cudaSuccess)
    return PLT_ERR_CUDA;
  k_chol_solve<<<static_cast<unsigned>(n_rhs), kSolveThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      factor, 0, n, nullptr, 0, rhs, out);
  return cudaGetLastError() == cudaSuccess ? PLT_OK : PLT_ERR_CUDA;
}

int plt_chol_solve_batched(const double* factor, int64_t n_batch, int n, const double* q_top, int l, const double* vals,
                           double* lam, void* stream) {
  if (!factor || !vals || !lam || n_batch < 0 || n < 1 || l < 0 || (l > 0 && !q_top)) return PLT_ERR_INVALID;
  if (!is_dev(factor) || !is_dev(vals) || !is_dev(lam)) return PLT_ERR_INVALID;
  if (n_batch == 0) return PLT_OK;
  const size_t smem = sizeof(double) * (((n + 31) & ~31) + kNB * (kNB + 1) + kNB);
  if (smem > 40 * 1024 &&
      cudaFuncSetAttribute((const void*)k_chol_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) !=
          cudaSuccess)
    return PLT_ERR_CUDA;
  k_chol_solve<<<static_cast<unsigned>(n_batch), kSolveThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      factor, static_cast<size_t>(n) * n, n, q_top, l, vals, lam);
  return cudaGetLastError() == cudaSuccess ? PLT_OK : PLT_ERR_CUDA;
}

}  // extern "C"
