// FMM operator kernels that do not depend on the RBF: P2M, M2M, multipole DFT, the
// Fourier-space M2L (Hadamard + inverse DFT), L2L, L2P, and the point pre/post passes.
// (RBF-dependent kernels -- P2P and the M2L operator tabulation -- are in fmm_rbf_ops.cu.)
//
// Layouts (all FP64):
//   M, L   : [cell][k][P]      P = order^dim nodes, node index row-major (axis 0 slowest)
//   Mhat   : [cell][km][F]     F = nf^(dim-1) * order half-spectrum, double2 (re, im)
//   Khat   : [7^dim offsets][kn][km][F]
//   points : SoA [dim][n]; weights SoA [km][n]; raw outputs SoA [kn][n]
#include "fmm_ops.cuh"

#include <cstdlib>

namespace plt {
namespace {

constexpr int kBlock = 128;
#ifndef PLT_HAD_WARPS3
#define PLT_HAD_WARPS3 24
#endif
#ifndef PLT_HAD_G3
#define PLT_HAD_G3 2
#endif

struct Mat3 {
  double a[9];
};

// ------------------------------------------------------------------------------------
// Point pre/post processing
// ------------------------------------------------------------------------------------
__global__ void k_transform_points(int dim, Mat3 A, const double* __restrict__ pts, int64_t n, int64_t ld,
                                   double* __restrict__ pos) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  double p[3] = {0, 0, 0};
  for (int b = 0; b < dim; ++b) p[b] = pts[i * dim + b];
  for (int a = 0; a < dim; ++a) {
    // geometry/point3d.hpp:36-40: p * A^T  => component a = sum_b A[a][b] p[b]
    double s = 0.0;
    for (int b = 0; b < dim; ++b) s += p[b] * A.a[a * dim + b];
    pos[a * ld + i] = s;
  }
}

__global__ void k_prepare_weights(int km, int fold, int dim, Mat3 A, const double* __restrict__ w,
                                  const int* __restrict__ perm, int64_t n, double* __restrict__ wt) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  int64_t j = perm ? perm[i] : i;
  if (!fold) {
    for (int m = 0; m < km; ++m) wt[m * n + i] = w[j * km + m];
  } else {
    // w' = A w  (gradient_kernel.hpp:53: grad_iso * A contracted with w)
    double v[3] = {0, 0, 0};
    for (int b = 0; b < dim; ++b) v[b] = w[j * dim + b];
    for (int a = 0; a < dim; ++a) {
      double s = 0.0;
      for (int b = 0; b < dim; ++b) s += A.a[a * dim + b] * v[b];
      wt[a * n + i] = s;
    }
  }
}

__global__ void k_finish_outputs(int kn, int fold, int dim, Mat3 A, const double* __restrict__ vt,
                                 const int* __restrict__ perm, int64_t n, int64_t lo, int64_t hi,
                                 double* __restrict__ out) {
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  int64_t j = perm ? perm[i] : i;
  const bool in = i >= lo && i < hi;
  if (!fold) {
    for (int b = 0; b < kn; ++b) out[j * kn + b] = in ? vt[b * n + i] : 0.0;
  } else {
    // out = A^T v
    double v[3] = {0, 0, 0};
    for (int a = 0; a < dim; ++a) v[a] = in ? vt[a * n + i] : 0.0;
    for (int b = 0; b < dim; ++b) {
      double s = 0.0;
      for (int a = 0; a < dim; ++a) s += A.a[a * dim + b] * v[a];
      out[j * dim + b] = s;
    }
  }
}

Mat3 make_mat3(int dim, const double* aniso) {
  Mat3 m{};
  for (int i = 0; i < dim * dim; ++i) m.a[i] = aniso[i];
  return m;
}

// ------------------------------------------------------------------------------------
// Interpolation helpers
// ------------------------------------------------------------------------------------
__device__ __forceinline__ double node_pos(int i, int order) { return -1.0 + 2.0 * i / (order - 1); }

// Barycentric basis values S_i(t), i < order, written with the given stride.
__device__ void bary_basis(int order, const double* __restrict__ beta, double t, double* s) {
  int hit = -1;
  double sum = 0.0;
  for (int i = 0; i < order; ++i) {
    double dt = t - node_pos(i, order);
    if (dt == 0.0) hit = i;
    double q = beta[i] / dt;
    s[i] = q;
    sum += q;
  }
  if (hit >= 0) {
    for (int i = 0; i < order; ++i) s[i] = i == hit ? 1.0 : 0.0;
  } else {
    double inv = 1.0 / sum;
    for (int i = 0; i < order; ++i) s[i] *= inv;
  }
}

template <int DIM>
__device__ __forceinline__ void cell_center(const Box& box, int level, uint32_t key, double (&c)[DIM],
                                            double& half) {
  int ci[DIM];
  morton_decode<DIM>(key, ci);
  const double w = box.width / static_cast<double>(1 << level);
  half = 0.5 * w;
#pragma unroll
  for (int a = 0; a < DIM; ++a) c[a] = box.center[a] - 0.5 * box.width + (ci[a] + 0.5) * w;
}

template <int DIM>
__device__ __forceinline__ void node_decode(int n, int order, int (&ni)[DIM]) {
#pragma unroll
  for (int a = DIM - 1; a >= 0; --a) {
    ni[a] = n % order;
    n /= order;
  }
}

// ------------------------------------------------------------------------------------
// P2M: M_c[m][n] = sum_j S_n(x_j) w_j[m]          (one CTA per source leaf)
// ------------------------------------------------------------------------------------
constexpr int kPointBatch = 32;

template <int DIM, int ORDER>
__global__ void __launch_bounds__(kBlock) k_p2m(TreeView tr, Box box, InterpDev it, int km,
                                                const double* __restrict__ wt, double* __restrict__ M) {
  extern __shared__ double sm[];
  const int p = ORDER > 0 ? ORDER : it.order;
  double* s_beta = sm;                          // [p]
  double* s_basis = s_beta + p;                 // [batch][DIM][p]
  double* s_w = s_basis + kPointBatch * DIM * p;  // [batch][km]
  const int leaf = tr.height - 1;
  const int cell = blockIdx.x;
  if (tr.flags && !(tr.flags[tr.cell_off[leaf] + cell] & kCellFlagM)) return;  // another rank's leaf
  const int tid = threadIdx.x;
  for (int i = tid; i < p; i += kBlock) s_beta[i] = it.beta[i];
  double c[DIM], half;
  cell_center<DIM>(box, leaf, tr.keys[tr.cell_off[leaf] + cell], c, half);
  const double inv_half = 1.0 / half;
  const int i0 = tr.leaf_start[cell], i1 = tr.leaf_start[cell + 1];
  int P = 1;
  for (int a = 0; a < DIM; ++a) P *= p;
  double* Mc = M + static_cast<size_t>(tr.cell_off[leaf] + cell) * km * P;
  __syncthreads();
  for (int b0 = i0; b0 < i1; b0 += kPointBatch) {
    const int nb = min(kPointBatch, i1 - b0);
    for (int e = tid; e < nb * DIM; e += kBlock) {
      int j = e / DIM, a = e % DIM;
      double t = (tr.pos[a * tr.n + b0 + j] - c[a]) * inv_half;
      bary_basis(p, s_beta, t, s_basis + (j * DIM + a) * p);
    }
    for (int e = tid; e < nb * km; e += kBlock) {
      int j = e / km, m = e % km;
      s_w[j * km + m] = wt[m * tr.n + b0 + j];
    }
    __syncthreads();
    for (int n = tid; n < P; n += kBlock) {
      int ni[DIM];
      node_decode<DIM>(n, p, ni);
      for (int m = 0; m < km; ++m) {
        double acc = b0 == i0 ? 0.0 : Mc[m * P + n];
        for (int j = 0; j < nb; ++j) {
          double s = s_w[j * km + m];
#pragma unroll
          for (int a = 0; a < DIM; ++a) s *= s_basis[(j * DIM + a) * p + ni[a]];
          acc += s;
        }
        Mc[m * P + n] = acc;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------
// Per-axis p x p contraction used by M2M and L2L.
//   out[o][r][i] = sum_q T(r, q) in[o][q][i],  array viewed as [outer][p][inner]
//   T(r, q) = tm[r * p + q] (transpose == false) or tm[q * p + r] (transpose == true)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void axis_contract(const double* __restrict__ in, double* out, int outer, int p,
                                              int inner, const double* __restrict__ tm, bool transpose,
                                              bool accumulate) {
  const int total = outer * p * inner;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    int i = e % inner;
    int r = (e / inner) % p;
    int o = e / (inner * p);
    const double* src = in + static_cast<size_t>(o) * p * inner + i;
    double acc = 0.0;
    if (!transpose) {
      for (int q = 0; q < p; ++q) acc += tm[r * p + q] * src[q * inner];
    } else {
      for (int q = 0; q < p; ++q) acc += tm[q * p + r] * src[q * inner];
    }
    if (accumulate) out[e] += acc; else out[e] = acc;
  }
}

// M2M: M_parent[m] = sum_children sum_n S_m^parent(y_n^child) M_child[n]   (CTA per parent cell)
template <int DIM, int ORDER>
__global__ void __launch_bounds__(kBlock) k_m2m(TreeView tr, int level, InterpDev it, int km,
                                                double* __restrict__ M) {
  extern __shared__ double sm[];
  const int p = ORDER > 0 ? ORDER : it.order;
  int P = 1;
  for (int a = 0; a < DIM; ++a) P *= p;
  double* s_t = sm;               // [2][p][p]
  double* s_acc = s_t + 2 * p * p;  // [P]
  double* s_b0 = s_acc + P;
  double* s_b1 = s_b0 + P;
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * p * p; i += kBlock) s_t[i] = it.child[i];
  const int cell = blockIdx.x;
  if (tr.flags && !(tr.flags[tr.cell_off[level] + cell] & kCellFlagM)) return;
  const uint32_t key = tr.keys[tr.cell_off[level] + cell];
  const int* dense_child = tr.dense + tr.dense_off[level + 1];
  double* Mp = M + static_cast<size_t>(tr.cell_off[level] + cell) * km * P;
  for (int m = 0; m < km; ++m) {
    for (int n = tid; n < P; n += kBlock) s_acc[n] = 0.0;
    for (int ch = 0; ch < (1 << DIM); ++ch) {
      int cidx = dense_child[(key << DIM) | ch];
      if (cidx < 0) continue;
      const double* Mc = M + (static_cast<size_t>(tr.cell_off[level + 1] + cidx) * km + m) * P;
      __syncthreads();
      for (int n = tid; n < P; n += kBlock) s_b0[n] = Mc[n];
      __syncthreads();
      double* in = s_b0;
      double* out = s_b1;
      int outer = 1, inner = P / p;
      for (int a = 0; a < DIM; ++a) {
        int side = (ch >> (DIM - 1 - a)) & 1;
        bool last = a == DIM - 1;
        // parent node r <- child node q: T(r, q) = child[side][r][q]
        axis_contract(in, last ? s_acc : out, outer, p, inner, s_t + side * p * p, false, last);
        __syncthreads();
        double* t = in; in = out; out = t;
        outer *= p;
        inner /= p;
      }
    }
    __syncthreads();
    for (int n = tid; n < P; n += kBlock) Mp[m * P + n] = s_acc[n];
    __syncthreads();
  }
}

// L2L: L_child[n] += sum_m S_m^parent(x_n^child) L_parent[m]   (CTA per child cell)
template <int DIM, int ORDER>
__global__ void __launch_bounds__(kBlock) k_l2l(TreeView tr, int level, InterpDev it, int kn,
                                                double* __restrict__ L, int cell_lo) {
  extern __shared__ double sm[];
  const int p = ORDER > 0 ? ORDER : it.order;
  int P = 1;
  for (int a = 0; a < DIM; ++a) P *= p;
  double* s_t = sm;
  double* s_b0 = s_t + 2 * p * p;
  double* s_b1 = s_b0 + P;
  const int tid = threadIdx.x;
  for (int i = tid; i < 2 * p * p; i += kBlock) s_t[i] = it.child[i];
  const int cell = cell_lo + blockIdx.x;
  const uint32_t key = tr.keys[tr.cell_off[level] + cell];
  const int pidx = tr.dense[tr.dense_off[level - 1] + (key >> DIM)];
  const int ch = key & ((1 << DIM) - 1);
  for (int b = 0; b < kn; ++b) {
    const double* Lp = L + (static_cast<size_t>(tr.cell_off[level - 1] + pidx) * kn + b) * P;
    double* Lc = L + (static_cast<size_t>(tr.cell_off[level] + cell) * kn + b) * P;
    __syncthreads();
    for (int n = tid; n < P; n += kBlock) s_b0[n] = Lp[n];
    __syncthreads();
    double* in = s_b0;
    double* out = s_b1;
    int outer = 1, inner = P / p;
    for (int a = 0; a < DIM; ++a) {
      int side = (ch >> (DIM - 1 - a)) & 1;
      bool last = a == DIM - 1;
      // child node r <- parent node q: T(r, q) = child[side][q][r]
      axis_contract(in, last ? Lc : out, outer, p, inner, s_t + side * p * p, true, last);
      __syncthreads();
      double* t = in; in = out; out = t;
      outer *= p;
      inner /= p;
    }
  }
}

// ------------------------------------------------------------------------------------
// L2P: v_i[b] = sum_n S_n(x_i) L_c[b][n]     (CTA per target leaf, warp per point)
// ------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(kBlock) k_l2p(TreeView tr, Box box, InterpDev it, int kn,
                                                const double* __restrict__ L, double* __restrict__ vt,
                                                int leaf_lo) {
  extern __shared__ double sm[];
  const int p = it.order;
  double* s_beta = sm;           // [p]
  double* s_basis = s_beta + p;  // [warps][DIM][p]
  const int leaf = tr.height - 1;
  const int cell = leaf_lo + blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < p; i += kBlock) s_beta[i] = it.beta[i];
  double c[DIM], half;
  cell_center<DIM>(box, leaf, tr.keys[tr.cell_off[leaf] + cell], c, half);
  const double inv_half = 1.0 / half;
  const int i0 = tr.leaf_start[cell], i1 = tr.leaf_start[cell + 1];
  int P = 1;
  for (int a = 0; a < DIM; ++a) P *= p;
  const double* Lc = L + static_cast<size_t>(tr.cell_off[leaf] + cell) * kn * P;
  double* basis = s_basis + warp * DIM * p;
  __syncthreads();
  for (int i = i0 + warp; i < i1; i += kBlock / 32) {
    if (lane < DIM) {
      double t = (tr.pos[lane * tr.n + i] - c[lane]) * inv_half;
      bary_basis(p, s_beta, t, basis + lane * p);
    }
    __syncwarp();
    double acc[kMaxDim] = {0, 0, 0};
    for (int n = lane; n < P; n += 32) {
      int ni[DIM];
      node_decode<DIM>(n, p, ni);
      double s = 1.0;
#pragma unroll
      for (int a = 0; a < DIM; ++a) s *= basis[a * p + ni[a]];
      for (int b = 0; b < kn; ++b) acc[b] += s * Lc[b * P + n];
    }
    for (int b = 0; b < kn; ++b) {
      double v = acc[b];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) vt[b * tr.n + i] = v;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------
// Fused last level of the downward pass (one CTA per parent cell of level leaf-1):
//   L_child = L2L(L_parent) + Lc[slot][child]   (M2L result of the leaf level, if any)
//   v_i     = L2P(L_child) for the points of the child
// The leaf-level expansions (the largest array of the whole evaluation: 8 P bytes x kn per
// leaf) live only in shared memory.
//   * L2L: per-axis contractions, register-blocked by columns (a thread loads the p inputs of a
//     column and produces its p outputs; the transfer matrices are kernel-parameter constants).
//     The first dim-1 axes are shared between siblings (axis 0: 2 variants, axis 1: 4), the last
//     axis produces all 2^dim children.
//   * L2P: warp <-> child, lane <-> point; the lane builds its 1-D bases in registers and
//     contracts the child's tensor axis by axis (p^dim + p^(dim-1) + ... FMAs), every lane of the
//     warp reading the same expansion entries (shared-memory broadcast).
//     The barycentric basis is evaluated in product form,
//       S_m(t) = beta_m prod_{j != m}(t - t_j) / sum_i beta_i prod_{j != i}(t - t_j),
//     algebraically the quotient form of interp.hpp (both numerator and denominator multiplied
//     by prod_j (t - t_j)) without its divisions and without the t == t_m special case.
// Shared layout: [children 2^dim x P][last shared stage]; the parent and the first stage are
// dead by the time the children are written and overlay the children area.
// ------------------------------------------------------------------------------------
// Threads per CTA by order: the shared-memory footprint (12 P doubles in 3-D) allows one resident CTA
// per SM from order 10 on, so the CTA itself has to bring the warps.
__host__ __device__ constexpr int leaf_threads(int order) { return order >= 10 ? 512 : (order >= 8 ? 256 : 128); }
constexpr int kLeafMaxOrder = 12;
constexpr int kLeafPosCap = 160;  // points of one parent staged in shared memory (else read from HBM)

struct LeafTables {
  double child[2 * kLeafMaxOrder * kLeafMaxOrder];  // [2][p][p]
  double beta[kLeafMaxOrder];
};

template <int DIM>
__host__ __device__ constexpr int leaf_stage_cells() {
  return (1 << DIM) + (DIM == 1 ? 1 : (DIM == 2 ? 2 : 4));
}

// out[r * stride] = sum_q T[q][r] in[q * stride]   (child node r <- parent node q)
template <int p>
__device__ __forceinline__ void leaf_contract_col(const double* __restrict__ tm, const double* in, int stride,
                                                  double* out) {
  double x[p];
#pragma unroll
  for (int q = 0; q < p; ++q) x[q] = in[q * stride];
#pragma unroll
  for (int r = 0; r < p; ++r) {
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < p; ++q) acc += tm[q * p + r] * x[q];
    out[r * stride] = acc;
  }
}

template <int p>
__device__ __forceinline__ void product_basis(const double* __restrict__ beta, double t, double (&s)[p]) {
  double d[p], pre[p];
#pragma unroll
  for (int m = 0; m < p; ++m) d[m] = t - (-1.0 + 2.0 * m / (p - 1));
  pre[0] = 1.0;
#pragma unroll
  for (int m = 1; m < p; ++m) pre[m] = pre[m - 1] * d[m - 1];
  double suf = 1.0, sum = 0.0;
#pragma unroll
  for (int m = p - 1; m >= 0; --m) {
    s[m] = beta[m] * (pre[m] * suf);
    sum += s[m];
    suf *= d[m];
  }
  const double inv = 1.0 / sum;
#pragma unroll
  for (int m = 0; m < p; ++m) s[m] *= inv;
}

template <int DIM, int ORDER>
__global__ void __launch_bounds__(leaf_threads(ORDER), ORDER <= 6 ? 6 : (ORDER <= 8 ? 2 : 1)) k_l2l_l2p_leaf(TreeView tr, Box box, LeafTables tb, int kn,
                                                               const double* __restrict__ L,
                                                               const double* __restrict__ Lc,
                                                               const int* __restrict__ leaf_meta,
                                                               double* __restrict__ vt, int par_lo, int leaf_lo,
                                                               int leaf_hi) {
  extern __shared__ double sm[];
  constexpr int NC = 1 << DIM;
  constexpr int kLeafThreads = leaf_threads(ORDER);
  constexpr int NW = kLeafThreads / 32;
  constexpr int p = ORDER;
  constexpr int P = DIM == 1 ? p : (DIM == 2 ? p * p : p * p * p);
  constexpr int PC = P / p;  // columns per axis
  // 3-D: the L2P lanes of a point read p different i0-slabs of the child at once; a slab stride of
  // p^2 doubles maps them all to the same bank (p even), p^2 + 1 spreads them over p banks.
  constexpr int S = DIM == 3 ? p * p + 1 : PC;  // stride of the leading index of a child
  constexpr int PS = DIM == 3 ? p * S : P;      // stride between children
  double* s_child = sm;                         // [NC][PS]
  double* s_last = s_child + NC * PS;           // last shared stage: [1 | 2 | 4][P]
  double* lvl0 = DIM == 1 ? s_last : s_child;   // parent
  double* lvl1 = DIM == 2 ? s_last : s_child + P;  // after axis 0 (DIM >= 2)
  __shared__ int s_first[NC];  // first point of the child, or -1
  __shared__ int s_count[NC];
  // Positions of the parent's points (the children are consecutive leaves, so their points are one
  // contiguous run of the sorted order): fetched once, coalesced, while the L2L stages run.
  __shared__ double s_pos[DIM * kLeafPosCap];
  __shared__ int s_run[2];     // first point and length of the run
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int leaf = tr.height - 1, pl = leaf - 1;
  const int pidx = par_lo + blockIdx.x;
  // one coalesced read of the parent's metadata (plan.cuh: leaf_meta) replaces key -> dense map -> leaf_start
  const int* meta = leaf_meta + static_cast<size_t>(pidx) * (2 + 3 * NC);
  const uint32_t pkey = static_cast<uint32_t>(meta[0]);
  const bool has_parent = L != nullptr;  // parent level >= 2
  const int slot = meta[1];

  if (tid < 32) {
    int first = -1, cnt = 0;
    if (tid < NC) {
      const int cidx = meta[2 + 3 * tid];
      if (cidx >= leaf_lo && cidx < leaf_hi) {  // also rejects -1
        first = meta[3 + 3 * tid];
        cnt = meta[4 + 3 * tid];
      }
      s_first[tid] = first;
      s_count[tid] = cnt;
    }
    int lo = first >= 0 ? first : 0x7fffffff, hi = first >= 0 ? first + cnt : 0;
    for (int o = 16; o > 0; o >>= 1) {
      lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (tid == 0) {
      s_run[0] = lo;
      s_run[1] = hi > lo ? hi - lo : 0;
    }
  }
  __syncthreads();
  const int pos_base = s_run[0];
  const bool pos_in_smem = s_run[1] <= kLeafPosCap;
  if (pos_in_smem) {
    for (int e = tid; e < DIM * s_run[1]; e += kLeafThreads) {
      const int a = e / s_run[1], q = e - a * s_run[1];
      s_pos[a * kLeafPosCap + q] = tr.pos[a * tr.n + pos_base + q];
    }
  }

  for (int b = 0; b < kn; ++b) {
    __syncthreads();  // tables ready / previous component done with the children
    if (has_parent) {
      const double* Lp = L + (static_cast<size_t>(tr.cell_off[pl] + pidx) * kn + b) * P;
      for (int n = tid; n < P; n += kLeafThreads) lvl0[n] = Lp[n];
      __syncthreads();
      if constexpr (DIM >= 2) {
        // axis 0 (stride P / p): parent -> 2 variants
        for (int item = tid; item < 2 * PC; item += kLeafThreads) {
          const int v = item / PC, col = item - v * PC;
          leaf_contract_col<p>(tb.child + v * p * p, lvl0 + col, PC, lvl1 + v * P + col);
        }
        __syncthreads();
      }
      if constexpr (DIM == 3) {
        // axis 1 (stride p): 2 variants -> 4.  column = (i, k)
        for (int item = tid; item < 4 * PC; item += kLeafThreads) {
          const int v = item / PC, col = item - v * PC;
          const int i = col / p, k = col - i * p;
          leaf_contract_col<p>(tb.child + (v & 1) * p * p, lvl1 + (v >> 1) * P + i * p * p + k, p,
                               s_last + v * P + i * p * p + k);
        }
        __syncthreads();
      }
    }
    // last axis (stride 1) -> all children; k_l2l accumulates the parent onto the M2L result
    for (int item = tid; item < NC * PC; item += kLeafThreads) {
      const int ch = item / PC, col = item - ch * PC;
      if (s_first[ch] < 0) continue;
      double* out = s_child + ch * PS + (DIM == 3 ? (col / p) * S + (col % p) * p : col * p);
      if (has_parent) {
        leaf_contract_col<p>(tb.child + (ch & 1) * p * p, s_last + (ch >> 1) * P + col * p, 1, out);
      } else {
#pragma unroll
        for (int r = 0; r < p; ++r) out[r] = 0.0;
      }
      if (slot >= 0) {
        const double* add = Lc + ((static_cast<size_t>(slot) * NC + ch) * kn + b) * P + col * p;
#pragma unroll
        for (int r = 0; r < p; ++r) out[r] = add[r] + out[r];
      }
    }
    __syncthreads();
    // L2P
    if constexpr (DIM == 3) {
      // p lanes per point (lane <-> slab i0 of the child's tensor), PP = 32 / p points per warp pass;
      // the (child, point group) passes of the whole parent are dealt round-robin to the warps.
      constexpr int PP = 32 / p;
      const int sub = lane / p, i0 = lane - sub * p;
      int pre[NC + 1];
      pre[0] = 0;
#pragma unroll
      for (int c = 0; c < NC; ++c) pre[c + 1] = pre[c] + (s_first[c] >= 0 ? (s_count[c] + PP - 1) / PP : 0);
      for (int g = warp; g < pre[NC]; g += NW) {
        int ch = 0;
#pragma unroll
        for (int c = 1; c < NC; ++c)
          if (g >= pre[c]) ch = c;
        int gbase = 0;
#pragma unroll
        for (int c = 0; c < NC; ++c)
          if (c == ch) gbase = pre[c];
        const int first = s_first[ch], cnt = s_count[ch];
        double c[DIM], half;
        cell_center<DIM>(box, leaf, (pkey << DIM) | ch, c, half);
        const double inv_half = 1.0 / half;
        const double* Lch = s_child + ch * PS;
        const int j = (g - gbase) * PP + sub;
        const bool act = sub < PP && j < cnt;
        const int i = first + (act ? j : 0);
        // Barycentric basis in product form, shared by the p lanes of a point: lane i0 computes only
        // its own un-normalised weights u_a = beta[i0] * prod_{m != i0} (t_a - x_m) on the three axes,
        // the lanes exchange u_1, u_2 with shuffles, and the three normalisations 1 / sum_m u_a[m]
        // are applied once to the point's total.
        double u[DIM];
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
          const double t = ((pos_in_smem ? s_pos[a * kLeafPosCap + (i - pos_base)] : tr.pos[a * tr.n + i]) - c[a]) * inv_half;
          double prod = tb.beta[0];
#pragma unroll
          for (int m = 1; m < p; ++m)
            if (i0 == m) prod = tb.beta[m];
#pragma unroll
          for (int m = 0; m < p; ++m) {
            const double dm = t - (-1.0 + 2.0 * m / (p - 1));
            prod *= (i0 == m) ? 1.0 : dm;
          }
          u[a] = prod;
        }
        const int lane0 = sub * p;  // first lane of the point (lanes beyond PP * p idle along)
        double u1[p], u2[p], s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int m = 0; m < p; ++m) {
          u1[m] = __shfl_sync(0xffffffffu, u[1], (lane0 + m) & 31);
          u2[m] = __shfl_sync(0xffffffffu, u[2], (lane0 + m) & 31);
          s1 += u1[m];
          s2 += u2[m];
        }
        const double* Ls = Lch + i0 * S;
        double ri = 0.0;
#pragma unroll
        for (int i1 = 0; i1 < p; ++i1) {
          double r = 0.0;
#pragma unroll
          for (int k = 0; k < p; ++k) r = fma(u2[k], Ls[i1 * p + k], r);
          ri = fma(u1[i1], r, ri);
        }
        const double v = u[0] * ri;
        // sums over the p lanes of the point; the point's first lane ends up with the totals
        double tot = v, s0 = u[0];
#pragma unroll
        for (int m = 1; m < p; ++m) {
          tot += __shfl_down_sync(0xffffffffu, v, m);
          s0 += __shfl_down_sync(0xffffffffu, u[0], m);
        }
        if (act && i0 == 0) vt[b * tr.n + i] = tot / (s0 * s1 * s2);
      }
    } else {
      for (int ch = warp; ch < NC; ch += NW) {
        const int first = s_first[ch], cnt = s_count[ch];
        if (first < 0) continue;
        double c[DIM], half;
        cell_center<DIM>(box, leaf, (pkey << DIM) | ch, c, half);
        const double inv_half = 1.0 / half;
        const double* Lch = s_child + ch * P;
        for (int j = lane; j < cnt; j += 32) {
          const int i = first + j;
          double bs[DIM][p];
#pragma unroll
          for (int a = 0; a < DIM; ++a)
            product_basis<p>(tb.beta, (tr.pos[a * tr.n + i] - c[a]) * inv_half, bs[a]);
          double v = 0.0;
          if constexpr (DIM == 1) {
#pragma unroll
            for (int k = 0; k < p; ++k) v = fma(bs[0][k], Lch[k], v);
          } else {
#pragma unroll
            for (int i0 = 0; i0 < p; ++i0) {
              double r = 0.0;
#pragma unroll
              for (int k = 0; k < p; ++k) r = fma(bs[1][k], Lch[i0 * p + k], r);
              v = fma(bs[0][i0], r, v);
            }
          }
          vt[b * tr.n + i] = v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// L2L of one level, 3-D, one CTA per PARENT cell: the first two axis contractions are shared between siblings
// (axis 0: 2 variants, axis 1: 4), the last one produces all 2^dim children and adds them onto the children's locals
// in HBM (which hold their own M2L result).  Register-blocked columns with the transfer matrices as kernel-parameter
// constants, as in the fused leaf pass: 2 p^4 (1 + 2 + 4) FMAs per parent instead of 8 x 3 p^4 for child-wise CTAs,
// 3 barriers per parent instead of 4 per child.
// ------------------------------------------------------------------------------------
template <int ORDER>
__global__ void __launch_bounds__(leaf_threads(ORDER)) k_l2l_parent3(TreeView tr, int level, LeafTables tb, int kn,
                                                                     double* __restrict__ L, int par_lo, int cell_lo,
                                                                     int cell_hi) {
  extern __shared__ double sm[];
  constexpr int DIM = 3, NC = 8, p = ORDER, P = p * p * p, PC = p * p;
  constexpr int NT = leaf_threads(ORDER);
  double* lvl0 = sm;            // parent [P]
  double* lvl1 = lvl0 + P;      // after axis 0: [2][P]
  double* lvl2 = lvl1 + 2 * P;  // after axis 1: [4][P]
  __shared__ int s_child[NC];
  const int tid = threadIdx.x;
  const int pl = level - 1;
  const int pidx = par_lo + blockIdx.x;
  const uint32_t pkey = tr.keys[tr.cell_off[pl] + pidx];
  if (tid < NC) {
    const int cidx = tr.dense[tr.dense_off[level] + ((pkey << DIM) | tid)];
    s_child[tid] = (cidx >= cell_lo && cidx < cell_hi) ? cidx : -1;
  }
  for (int b = 0; b < kn; ++b) {
    const double* Lp = L + (static_cast<size_t>(tr.cell_off[pl] + pidx) * kn + b) * P;
    __syncthreads();  // previous component done with the stages; s_child written
    for (int n = tid; n < P; n += NT) lvl0[n] = Lp[n];
    __syncthreads();
    for (int item = tid; item < 2 * PC; item += NT) {  // axis 0 (stride p^2)
      const int v = item / PC, col = item - v * PC;
      leaf_contract_col<p>(tb.child + v * p * p, lvl0 + col, PC, lvl1 + v * P + col);
    }
    __syncthreads();
    for (int item = tid; item < 4 * PC; item += NT) {  // axis 1 (stride p): column = (i, k)
      const int v = item / PC, col = item - v * PC;
      const int i = col / p, k = col - i * p;
      leaf_contract_col<p>(tb.child + (v & 1) * p * p, lvl1 + (v >> 1) * P + i * p * p + k, p, lvl2 + v * P + i * p * p + k);
    }
    __syncthreads();
    for (int item = tid; item < NC * PC; item += NT) {  // axis 2 (stride 1) -> the children, accumulated in HBM
      const int ch = item / PC, col = item - ch * PC;
      const int cidx = s_child[ch];
      if (cidx < 0) continue;
      double out[p];
      leaf_contract_col<p>(tb.child + (ch & 1) * p * p, lvl2 + (ch >> 1) * P + col * p, 1, out);
      double* Lc = L + (static_cast<size_t>(tr.cell_off[level] + cidx) * kn + b) * P + col * p;
#pragma unroll
      for (int r = 0; r < p; ++r) Lc[r] += out[r];
    }
  }
}

// ------------------------------------------------------------------------------------
// Leaf pass for the POLYNOMIAL interpolant (d = kClassic: orders 6, 8, 10), 3-D, without the last L2L.
//
// The parent's local expansion is the tensor polynomial F(x) = sum_m L_parent[m] S_m^parent(x) of degree
// order - 1 per axis.  L2L samples it at the child's nodes and L2P interpolates those samples with the child's
// basis -- again a polynomial of degree order - 1 through `order` samples per axis of a polynomial of that same
// degree, i.e. F itself.  So
//     v(x) = sum_m S_m^parent(x) L_parent[m]  +  sum_n S_n^child(x) Lc_child[n]
// (Lc = the child's own M2L result) is the same function as L2P(L2L(L_parent) + Lc), evaluated without forming the
// 2^dim child expansions: two tensor contractions per point and no CTA-wide barriers between dependent stages
// (the fused kernel above: 3 contraction stages + 5 barriers per parent, 7 100 warp instructions per parent at
// order 6, latency bound at 0.06 of HBM; this one: ~2 000).  Differences to the staged form are rounding only.
//
// One CTA per parent cell of level leaf-1:
//   * warp 0 issues one TMA bulk copy (cp.async.bulk -> mbarrier) per i0-slab of the parent expansion and of the
//     present children's Lc into a padded shared layout (slab stride p^2 + 2 doubles: the p lanes of a point read
//     p different slabs, which must fall into different banks);
//   * meanwhile all threads evaluate the normalised 1-D barycentric bases of the parent's points, one
//     (point, axis) task per thread, in the child frame t and in the parent frame (t +- 1) / 2, product form
//     (no special case at a node), into shared memory;
//   * after the mbarrier flips: p lanes per point (lane <-> slab i0), 32 / p points per warp pass; each lane
//     contracts its slab of both expansions with the point's bases, the p partial sums are added by shuffles.
// ------------------------------------------------------------------------------------
constexpr int kLeafDirectThreads = 128;
constexpr int kLeafDirectCap = 64;  // points whose bases are staged at a time

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on the mbarrier (16-byte aligned, size % 16 == 0).
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Normalised barycentric basis S_m(t), m < p, product form (no division by t - x_m).
template <int p>
__device__ __forceinline__ void product_basis_store(const double* __restrict__ beta, double t, double* out) {
  double d[p], pre[p], s[p];
#pragma unroll
  for (int m = 0; m < p; ++m) d[m] = t - (-1.0 + 2.0 * m / (p - 1));
  pre[0] = 1.0;
#pragma unroll
  for (int m = 1; m < p; ++m) pre[m] = pre[m - 1] * d[m - 1];
  double suf = 1.0, sum = 0.0;
#pragma unroll
  for (int m = p - 1; m >= 0; --m) {
    s[m] = beta[m] * (pre[m] * suf);
    sum += s[m];
    suf *= d[m];
  }
  const double inv = 1.0 / sum;
#pragma unroll
  for (int m = 0; m < p; ++m) out[m] = s[m] * inv;
}

template <int ORDER>
__global__ void __launch_bounds__(kLeafDirectThreads) k_leaf_direct3(TreeView tr, Box box, LeafTables tb, int kn,
                                                                     const double* __restrict__ L,
                                                                     const double* __restrict__ Lc,
                                                                     const int* __restrict__ leaf_meta,
                                                                     double* __restrict__ vt, int par_lo, int leaf_lo,
                                                                     int leaf_hi) {
  constexpr int DIM = 3, NC = 8, p = ORDER, P = p * p * p;
  constexpr int S = p * p + 2;   // slab stride (doubles): 16-byte aligned, p slabs in p different banks
  constexpr int ES = p * S;      // expansion stride
  constexpr int PP = 32 / p;     // points per warp pass
  constexpr int NW = kLeafDirectThreads / 32;
  constexpr int BS = 2 * DIM * p;  // basis doubles per point: [frame: child, parent][axis][p]
  extern __shared__ __align__(16) double sm[];
  double* sL = sm;               // parent expansion [p][S]
  double* sC = sL + ES;          // children's own M2L results [NC][p][S]
  double* sB = sC + NC * ES;     // bases [kLeafDirectCap][BS]
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ int s_first[NC], s_count[NC], s_run[2];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int leaf = tr.height - 1, pl = leaf - 1;
  const int pidx = par_lo + blockIdx.x;
  const int* meta = leaf_meta + static_cast<size_t>(pidx) * (2 + 3 * NC);
  const uint32_t pkey = static_cast<uint32_t>(meta[0]);
  const int slot = meta[1];
  const bool has_parent = L != nullptr, has_own = slot >= 0;

  if (tid < 32) {
    int first = -1, cnt = 0;
    if (tid < NC) {
      const int cidx = meta[2 + 3 * tid];
      if (cidx >= leaf_lo && cidx < leaf_hi) {  // also rejects -1
        first = meta[3 + 3 * tid];
        cnt = meta[4 + 3 * tid];
      }
      s_first[tid] = first;
      s_count[tid] = cnt;
    }
    int lo = first >= 0 ? first : 0x7fffffff, hi = first >= 0 ? first + cnt : 0;
    for (int o = 16; o > 0; o >>= 1) {
      lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (tid == 0) {
      s_run[0] = lo;
      s_run[1] = hi > lo ? hi - lo : 0;
      mbar_init(&s_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  }
  __syncthreads();
  const int run0 = s_run[0], n_run = s_run[1];
  if (n_run == 0) return;  // no target of this shard below the parent (uniform)
  // the children are consecutive leaves: a point's child = the range of the sorted order it falls into
  int cfirst[NC], ccnt[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    cfirst[c] = s_first[c];
    ccnt[c] = s_count[c];
  }
  double pc[DIM], phalf;
  cell_center<DIM>(box, pl, pkey, pc, phalf);
  const double inv_half = 2.0 / phalf;  // 1 / (half width of a child)

  for (int b = 0; b < kn; ++b) {
    // ---- stage the expansions of component b (asynchronous; overlapped with the basis evaluation) ----
    uint32_t bytes = 0;
    if (has_parent) bytes += P * 8;
    if (has_own) {
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (cfirst[c] >= 0) bytes += P * 8;
    }
    if (b > 0) __syncthreads();  // the previous component's expansions are no longer read
    if (warp == 0 && bytes > 0) {
      if (lane == 0) mbar_expect_tx(&s_bar, bytes);
      __syncwarp();
      if (has_parent) {
        const double* src = L + (static_cast<size_t>(tr.cell_off[pl] + pidx) * kn + b) * P;
        if (lane < p) bulk_g2s(sL + lane * S, src + lane * p * p, p * p * 8, &s_bar);
      }
      if (has_own) {
        for (int e = lane; e < NC * p; e += 32) {
          const int c = e / p, i0 = e - c * p;
          if (s_first[c] < 0) continue;
          const double* src = Lc + ((static_cast<size_t>(slot) * NC + c) * kn + b) * P;
          bulk_g2s(sC + c * ES + i0 * S, src + i0 * p * p, p * p * 8, &s_bar);
        }
      }
    }

    for (int q0 = 0; q0 < n_run; q0 += kLeafDirectCap) {
      const int nq = min(kLeafDirectCap, n_run - q0);
      // ---- bases: one (point, axis) task per thread, both frames ----
      if (b == 0 || n_run > kLeafDirectCap) {
        if (q0 > 0 || b > 0) __syncthreads();  // previous chunk's bases no longer read
        for (int task = tid; task < nq * DIM; task += kLeafDirectThreads) {
          const int q = task / DIM, a = task - q * DIM;
          const int i = run0 + q0 + q;
          int ch = 0;
#pragma unroll
          for (int c = 1; c < NC; ++c)
            if (cfirst[c] >= 0 && i >= cfirst[c]) ch = c;  // children ascend with the sorted order
          const int side = (ch >> (DIM - 1 - a)) & 1;
          const double x = tr.pos[a * tr.n + i];
          // child frame: centre of the child = parent centre -+ half of the child's width
          const double cc = pc[a] + (side ? 0.5 : -0.5) * phalf;
          const double t = (x - cc) * inv_half;
          const double tp = (x - pc[a]) * (0.5 * inv_half);
          double* dst = sB + q * BS + a * p;
          product_basis_store<p>(tb.beta, t, dst);
          product_basis_store<p>(tb.beta, tp, dst + DIM * p);
        }
        __syncthreads();
      }
      if (q0 == 0 && bytes > 0) mbar_wait(&s_bar, b & 1);

      // ---- contraction: p lanes per point, PP points per warp pass ----
      const int sub = lane / p, i0 = lane - sub * p;
      for (int g = warp * PP; g < nq; g += NW * PP) {
        const int q = g + sub;
        const bool act = sub < PP && q < nq;
        const int qq = act ? q : g;
        const int i = run0 + q0 + qq;
        int ch = 0;
#pragma unroll
        for (int c = 1; c < NC; ++c)
          if (cfirst[c] >= 0 && i >= cfirst[c]) ch = c;
        const double* bs = sB + qq * BS;
        double val = 0.0;
#pragma unroll
        for (int fr = 0; fr < 2; ++fr) {
          if (fr == 0 ? !has_own : !has_parent) continue;  // uniform
          const double* E = (fr == 0 ? sC + ch * ES : sL) + i0 * S;
          const double* u = bs + fr * DIM * p;
          double u1[p], u2[p];
#pragma unroll
          for (int m = 0; m < p; m += 2) {
            const double2 a1 = *reinterpret_cast<const double2*>(u + p + m);
            const double2 a2 = *reinterpret_cast<const double2*>(u + 2 * p + m);
            u1[m] = a1.x; u1[m + 1] = a1.y;
            u2[m] = a2.x; u2[m + 1] = a2.y;
          }
          double acc = 0.0;
#pragma unroll
          for (int i1 = 0; i1 < p; ++i1) {
            double r = 0.0;
#pragma unroll
            for (int k = 0; k < p; k += 2) {
              const double2 e2 = *reinterpret_cast<const double2*>(E + i1 * p + k);
              r = fma(u2[k], e2.x, r);
              r = fma(u2[k + 1], e2.y, r);
            }
            acc = fma(u1[i1], r, acc);
          }
          val = fma(u[i0], acc, val);
        }
        // sum over the p lanes of the point; its first lane ends up with the total
        double tot = val;
#pragma unroll
        for (int m = 1; m < p; ++m) tot += __shfl_down_sync(0xffffffffu, val, m);
        if (act && i0 == 0) vt[b * tr.n + i] = tot;
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// Small dense DFT stages (generic pointers: shared or global).
//   out[o][k][i] = sum_n in[o][n][i] * W^(k n),   W = e^{-2 pi i / nf} (or its conjugate)
// ------------------------------------------------------------------------------------
template <bool IN_REAL>
__device__ __forceinline__ void dft_stage(const void* in_, double2* out, int outer, int n_in, int n_out,
                                          int inner, const double2* __restrict__ tw, int nf, bool conj) {
  const int total = outer * n_out * inner;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    int i = e % inner;
    int k = (e / inner) % n_out;
    int o = e / (inner * n_out);
    double re = 0.0, im = 0.0;
    int idx = 0;
    const size_t base = static_cast<size_t>(o) * n_in * inner + i;
    for (int n = 0; n < n_in; ++n) {
      double2 w = tw[idx];
      if (conj) w.y = -w.y;
      if constexpr (IN_REAL) {
        double v = static_cast<const double*>(in_)[base + static_cast<size_t>(n) * inner];
        re = fma(v, w.x, re);
        im = fma(v, w.y, im);
      } else {
        double2 v = static_cast<const double2*>(in_)[base + static_cast<size_t>(n) * inner];
        re = fma(v.x, w.x, re);
        re = fma(-v.y, w.y, re);
        im = fma(v.x, w.y, im);
        im = fma(v.y, w.x, im);
      }
      idx += k;
      if (idx >= nf) idx -= nf;
    }
    out[e] = make_double2(re, im);
  }
}

// Last inverse stage, half spectrum -> real: out[o][m] = Re in[o][0] + 2 sum_{k>=1} Re(in[o][k] e^{+i th k m})
__device__ __forceinline__ void idft_stage_c2r(const double2* in, double* out, int outer, int p,
                                               const double2* __restrict__ tw, int nf, bool accumulate) {
  const int total = outer * p;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    int m = e % p;
    int o = e / p;
    const double2* src = in + static_cast<size_t>(o) * p;
    double acc = src[0].x;
    int idx = m;
    for (int k = 1; k < p; ++k) {
      double2 w = tw[idx];  // (cos, -sin); Re(v e^{+i th}) = v.x cos - v.y sin = v.x w.x + v.y w.y
      acc = fma(2.0 * src[k].x, w.x, acc);
      acc = fma(2.0 * src[k].y, w.y, acc);
      idx += m;
      if (idx >= nf) idx -= nf;
    }
    if (accumulate) out[e] += acc; else out[e] = acc;
  }
}

// M -> Mhat for all cells of levels >= 2 (CTA grid-strides over cells).
// Shared (or global scratch) buffers: real P, complex bufA, complex bufB.
template <int DIM, int ORDER>
__global__ void __launch_bounds__(kBlock) k_m2hat(int first_cell, int n_cells, InterpDev it, int km,
                                                  const double* __restrict__ M, double2* __restrict__ Mhat,
                                                  double2* gscratch, int scratch_elems,
                                                  const unsigned char* __restrict__ flags) {
  extern __shared__ double2 sm2[];
  const int p = ORDER > 0 ? ORDER : it.order, nf = ORDER > 0 ? 2 * ORDER - 1 : it.nf;
  int P = 1, F = p;
  for (int a = 0; a < DIM; ++a) P *= p;
  for (int a = 0; a + 1 < DIM; ++a) F *= nf;
  double2* s_tw = sm2;  // [nf]
  double2* bufA = gscratch ? gscratch + static_cast<size_t>(blockIdx.x) * 2 * scratch_elems : sm2 + nf;
  double2* bufB = bufA + scratch_elems;
  for (int i = threadIdx.x; i < nf; i += kBlock) s_tw[i] = it.tw[i];
  __syncthreads();
  for (int cell = blockIdx.x; cell < n_cells * km; cell += gridDim.x) {
    if (flags && !(flags[first_cell + cell / km] & kCellFlagMhat)) continue;  // uniform across the CTA
    const double* Mc = M + static_cast<size_t>(first_cell) * km * P + static_cast<size_t>(cell) * P;
    double2* out = Mhat + static_cast<size_t>(cell) * F;
    // last axis: real -> half spectrum
    if constexpr (DIM == 1) {
      dft_stage<true>(Mc, out, 1, p, p, 1, s_tw, nf, false);
    } else {
      dft_stage<true>(Mc, bufA, P / p, p, p, 1, s_tw, nf, false);
      __syncthreads();
      // remaining axes, last-1 down to 0
      double2* in = bufA;
      double2* ob = bufB;
      int outer = P / p / p;  // p^(axis)
      int inner = p;          // already transformed tail
      for (int a = DIM - 2; a >= 0; --a) {
        double2* dst = a == 0 ? out : ob;
        dft_stage<false>(in, dst, outer, p, nf, inner, s_tw, nf, false);
        __syncthreads();
        double2* t = in; in = ob; ob = t;
        inner *= nf;
        outer /= p;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------
// M2L, Fourier space.
// ------------------------------------------------------------------------------------
template <int DIM>
struct M2LGeom {
  static constexpr int NC = 1 << DIM;  // children per cell
  static constexpr int NN = DIM == 1 ? 3 : (DIM == 2 ? 9 : 27);
  static constexpr int NOFF = DIM == 1 ? 7 : (DIM == 2 ? 49 : 343);
  static constexpr int CENTER = (NN - 1) / 2;
};

__device__ __forceinline__ void cfma(double2& acc, const double2& k, const double2& m) {
  acc.x = fma(k.x, m.x, acc.x);
  acc.x = fma(-k.y, m.y, acc.x);
  acc.y = fma(k.x, m.y, acc.y);
  acc.y = fma(k.y, m.x, acc.y);
}

// Generic Hadamard accumulation (any kn x km), one CTA per active target parent, threads over
// frequencies:
//   Lhat[slot][ct][b][f] = sum_{cs in list(ct)} sum_a Khat[o(ct,cs)][b][a][f] * Mhat[cs][a][f]
template <int DIM, int KN, int KM>
__global__ void __launch_bounds__(256) k_m2l_hadamard(M2LArgs a, int F) {
  constexpr int NC = M2LGeom<DIM>::NC, NN = M2LGeom<DIM>::NN;
  __shared__ int s_src[NN * NC];  // Mhat cell index of the source child or -1
  const int slot = blockIdx.x;
  for (int e = threadIdx.x; e < NN * NC; e += blockDim.x) s_src[e] = a.src_ids[static_cast<size_t>(slot) * NN * NC + e];
  const unsigned tmask = a.trg_mask[slot];
  __syncthreads();

  for (int f = threadIdx.x; f < F; f += blockDim.x) {
    double2 acc[NC][KN];
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int b = 0; b < KN; ++b) acc[c][b] = make_double2(0.0, 0.0);
    for (int nb = 0; nb < NN; ++nb) {
      if (nb == M2LGeom<DIM>::CENTER) continue;
      int e3[DIM], r = nb;
#pragma unroll
      for (int d = DIM - 1; d >= 0; --d) {
        e3[d] = (r % 3) - 1;
        r /= 3;
      }
#pragma unroll
      for (int cs = 0; cs < NC; ++cs) {
        const int sid = s_src[nb * NC + cs];
        if (sid < 0) continue;
        double2 mh[KM];
#pragma unroll
        for (int m = 0; m < KM; ++m) mh[m] = a.Mhat[(static_cast<size_t>(sid) * KM + m) * F + f];
#pragma unroll
        for (int ct = 0; ct < NC; ++ct) {
          if (!((tmask >> ct) & 1u)) continue;
          // offset o = source child coord - target child coord, per axis in [-3, 3]
          int oi = 0, far = 0;
#pragma unroll
          for (int d = 0; d < DIM; ++d) {
            int o = 2 * e3[d] + ((cs >> (DIM - 1 - d)) & 1) - ((ct >> (DIM - 1 - d)) & 1);
            far |= (o > 1 || o < -1);
            oi = oi * 7 + (o + 3);
          }
          if (!far) continue;
          const double2* kh = a.Khat + static_cast<size_t>(oi) * KN * KM * F + f;
#pragma unroll
          for (int b = 0; b < KN; ++b)
#pragma unroll
            for (int m = 0; m < KM; ++m) cfma(acc[ct][b], kh[static_cast<size_t>(b * KM + m) * F], mh[m]);
        }
      }
    }
#pragma unroll
    for (int ct = 0; ct < NC; ++ct)
#pragma unroll
      for (int b = 0; b < KN; ++b)
        a.Lhat[((static_cast<size_t>(slot) * NC + ct) * KN + b) * F + f] = acc[ct][b];
  }
}

// Scalar (kn = km = 1) Hadamard accumulation, the hot kernel of the far field.
//
// A CTA owns a tile of kHadTF = 32 consecutive frequencies and keeps the operator slice
// Khat[all 7^dim offsets][tile] resident in shared memory (3-D: 343 x 32 x 16 B = 171.5 KiB),
// then streams over its share of the active target parents, one parent per warp at a time:
// lanes = the 32 frequencies of the tile, accumulators = the 2^dim target children (complex,
// registers).  The parent's 3^dim x 2^dim source-id table is read with coalesced loads and
// compacted with ballots, so the loop runs over *present* source cells only (surface clouds
// fill a fifth of the table); their Mhat rows are fetched kHadG at a time (512 B per warp and
// row, coalesced) before the arithmetic to keep several loads in flight per warp.  No CTA-wide
// barrier inside the main loop.
constexpr int kHadTF = 32;
// Warps per CTA (one persistent CTA per SM) and Mhat rows in flight per warp and pipeline stage.  3-D: the 171.5 KiB
// operator slice leaves room for one CTA per SM, so the CTA brings all the warps.  Measured on config #3
// (profiles/r02_d_hadamard_variants.md): 16 warps x 4-row groups (120 registers) 12.46 ms; 20 x 3: 11.49;
// 24 x 2 (80 registers): 11.43; 28 x 2: 11.45; 32 x 2 (64 registers, spills): 13.3.
template <int DIM> struct HadCfg { static constexpr int kWarps = 16, kG = 4; };
template <> struct HadCfg<3> { static constexpr int kWarps = PLT_HAD_WARPS3, kG = PLT_HAD_G3; };

template <int DIM, bool VEC>
__global__ void __launch_bounds__(HadCfg<DIM>::kWarps * 32, 1) k_m2l_hadamard_tiled(M2LArgs a, int F, int n_ftiles,
                                                                                     int no_tma) {
  constexpr int NC = M2LGeom<DIM>::NC, NN = M2LGeom<DIM>::NN, NOFF = M2LGeom<DIM>::NOFF;
  constexpr int kHadWarps = HadCfg<DIM>::kWarps, kHadG = HadCfg<DIM>::kG;
  constexpr int NE = NN * NC;              // entries of the source-id table
  constexpr int NL = NE - NC;              // list capacity: the centre block has no far pair
  constexpr int NCH = (NE + 31) / 32;      // 32-entry chunks
  extern __shared__ double2 sm2[];
  double2* Ks = sm2;                                      // [NOFF][TF]
  int2* s_meta = reinterpret_cast<int2*>(Ks + NOFF * kHadTF);  // [NE]: x = offset index base, y = far mask
  int2* s_list = s_meta + NE;                                   // [warps][NL]: (Mhat row, base | far mask << 16)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __shared__ uint64_t s_bar;
  uint32_t bar_phase = 0;
  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }

  for (int code = threadIdx.x; code < NE; code += blockDim.x) {
    const int nb = code / NC, cs = code % NC;
    int e3[DIM], r = nb;
#pragma unroll
    for (int d = DIM - 1; d >= 0; --d) {
      e3[d] = (r % 3) - 1;
      r /= 3;
    }
    int base = 0, mask = 0;
#pragma unroll
    for (int d = 0; d < DIM; ++d) base = base * 7 + (2 * e3[d] + ((cs >> (DIM - 1 - d)) & 1) + 3);
    for (int ct = 0; ct < NC; ++ct) {
      bool far = false;
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        const int o = 2 * e3[d] + ((cs >> (DIM - 1 - d)) & 1) - ((ct >> (DIM - 1 - d)) & 1);
        far = far || o > 1 || o < -1;
      }
      if (far) mask |= 1 << ct;
    }
    s_meta[code] = make_int2(base, mask);
  }

  // Balanced persistent schedule: the n_ftiles x n_active (frequency tile, parent) items, tile-major,
  // are cut into gridDim.x equal contiguous ranges; a range spans at most a few tiles, and the CTA
  // re-stages the operator slice when it crosses a tile boundary.
  const long long n_items = static_cast<long long>(n_ftiles) * a.n_active;
  const long long q_lo = n_items * blockIdx.x / gridDim.x, q_hi = n_items * (blockIdx.x + 1) / gridDim.x;
  int2* list = s_list + warp * NL;
  for (long long q0 = q_lo; q0 < q_hi;) {
    const int ftile = static_cast<int>(q0 / a.n_active);
    const int slot_lo = static_cast<int>(q0 - static_cast<long long>(ftile) * a.n_active);
    // a segment ends with the CTA's range, the tile or the level (another level = another operator slice)
    int lv = 0, lv_end = a.n_active;
    if (a.n_lvls > 1) {
      while (lv + 1 < a.n_lvls && slot_lo >= a.lvl_slot_end[lv]) ++lv;
      lv_end = a.lvl_slot_end[lv];
    }
    const double2* Khat_l = a.Khat + static_cast<size_t>(lv) * a.khat_level_stride;
    const long long seg_end = min(q_hi, static_cast<long long>(ftile) * a.n_active + lv_end);
    const int slot_hi = slot_lo + static_cast<int>(seg_end - q0);
    q0 = seg_end;
    const int f = ftile * kHadTF + lane;
    const bool fok = f < F;
    // Vector kinds (kn x km > 1): one (b, a) component pair at a time with its own operator slice;
    // the partial sums over a are carried through Lhat by the warp that owns the parent.
    const int kn = VEC ? a.kn : 1, km = VEC ? a.km : 1;  // scalar instantiation: compile-time 1 x 1
    for (int comp = 0; comp < kn * km; ++comp) {
    const int cb = comp / km, ca = comp - cb * km;
    __syncthreads();  // previous operator slice no longer in use (and s_meta written)
    // Stage the slice: NOFF rows of (up to) 32 frequencies = 512 contiguous bytes each, as TMA bulk copies counted on
    // an mbarrier (issued by warp 0, ~11 per lane; no staging registers, ~2 us instead of ~12 for 171.5 KiB), the tail of
    // a partial last tile zero-filled by the threads.
    if (!no_tma) {
      const int valid = min(kHadTF, F - ftile * kHadTF);
      if (warp == 0) {
        if (lane == 0) mbar_expect_tx(&s_bar, static_cast<uint32_t>(NOFF) * valid * sizeof(double2));
        __syncwarp();
        for (int oi = lane; oi < NOFF; oi += 32)
          bulk_g2s(Ks + oi * kHadTF, Khat_l + (static_cast<size_t>(oi) * kn * km + comp) * F + ftile * kHadTF,
                   valid * sizeof(double2), &s_bar);
      }
      if (valid < kHadTF)
        for (int e = threadIdx.x; e < NOFF * (kHadTF - valid); e += blockDim.x)
          Ks[(e / (kHadTF - valid)) * kHadTF + valid + e % (kHadTF - valid)] = make_double2(0.0, 0.0);
      mbar_wait(&s_bar, bar_phase);
      bar_phase ^= 1u;
    } else {
      for (int e = threadIdx.x; e < NOFF * kHadTF; e += blockDim.x) {
        const int oi = e / kHadTF, ff = ftile * kHadTF + (e % kHadTF);
        Ks[e] = ff < F ? Khat_l[(static_cast<size_t>(oi) * kn * km + comp) * F + ff] : make_double2(0.0, 0.0);
      }
    }
    __syncthreads();

    for (int slot = slot_lo + warp; slot < slot_hi; slot += kHadWarps) {
      const int* tab = a.src_ids + static_cast<size_t>(slot) * NE;
      const int tmask = a.trg_mask[slot];
      // compact the present source cells that have a far target child
      int n = 0;
      __syncwarp();  // previous parent done with the list
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int idx = c * 32 + lane;
        const int sid = idx < NE ? __ldcs(tab + idx) : -1;
        int2 meta = make_int2(0, 0);
        if (idx < NE) meta = s_meta[idx];
        const int fm = meta.y & tmask;
        const bool pres = sid >= 0 && fm != 0;
        const unsigned m = __ballot_sync(0xffffffffu, pres);
        if (pres) list[n + __popc(m & ((1u << lane) - 1u))] = make_int2(sid, meta.x | (fm << 16));
        n += __popc(m);
      }
      __syncwarp();
      double2 acc[NC];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        acc[c] = make_double2(0.0, 0.0);
        if (VEC && ca > 0 && fok && ((tmask >> c) & 1)) acc[c] = a.Lhat[((static_cast<size_t>(slot) * NC + c) * kn + cb) * F + f];
      }

      // software pipeline over groups of kHadG entries: the Mhat rows of the next group are in
      // flight while the current group is multiplied
      auto load = [&](double2 (&mh)[kHadG], int (&pk)[kHadG], int base) {
#pragma unroll
        for (int g = 0; g < kHadG; ++g) {
          const int e = base + g;
          int2 le = make_int2(0, 0);
          if (e < n) le = list[e];
          pk[g] = le.y;
          mh[g] = (e < n && fok) ? a.Mhat[(static_cast<size_t>(le.x) * km + ca) * F + f] : make_double2(0.0, 0.0);
        }
      };
      auto compute = [&](const double2 (&mh)[kHadG], const int (&pk)[kHadG], int base) {
#pragma unroll
        for (int g = 0; g < kHadG; ++g) {
          if (base + g >= n) break;  // warp-uniform
          const int fm = pk[g] >> 16;
          const double2* kp = Ks + (pk[g] & 0xffff) * kHadTF + lane;
#pragma unroll
          for (int ct = 0; ct < NC; ++ct) {
            int cto = 0;  // compile-time: sum_d ct_d 7^(DIM-1-d)
#pragma unroll
            for (int d = 0; d < DIM; ++d) cto = cto * 7 + ((ct >> (DIM - 1 - d)) & 1);
            if ((fm >> ct) & 1) cfma(acc[ct], kp[-cto * kHadTF], mh[g]);
          }
        }
      };
      double2 mhA[kHadG], mhB[kHadG];
      int pkA[kHadG], pkB[kHadG];
      load(mhA, pkA, 0);
      for (int g0 = 0; g0 < n; g0 += 2 * kHadG) {
        load(mhB, pkB, g0 + kHadG);
        compute(mhA, pkA, g0);
        load(mhA, pkA, g0 + 2 * kHadG);
        compute(mhB, pkB, g0 + kHadG);
      }
      if (fok) {
#pragma unroll
        for (int ct = 0; ct < NC; ++ct)
          if ((tmask >> ct) & 1)  // written once, read once by the IDFT: evict-first keeps the Mhat rows in L2
            __stcs(&a.Lhat[((static_cast<size_t>(slot) * NC + ct) * kn + cb) * F + f], acc[ct]);
      }
    }
    }
  }
}

// Level of a slot of a (possibly multi-level) chunk, M2LArgs::n_lvls / lvl_slot_end.
__device__ __forceinline__ int m2l_slot_level(const M2LArgs& a, int slot) {
  int lv = 0;
  while (lv + 1 < a.n_lvls && slot >= a.lvl_slot_end[lv]) ++lv;
  return a.level + lv;
}

// Inverse DFT of the accumulated spectra, pruned to the order^dim nodes:  L[cell][b][:] = IDFT(Lhat)
// One CTA per (slot, child, b).
template <int DIM, int ORDER>
__global__ void __launch_bounds__(kBlock) k_m2l_idft(M2LArgs a, InterpDev it, double2* gscratch,
                                                     int scratch_elems) {
  extern __shared__ double2 sm2[];
  constexpr int NC = 1 << DIM;
  const int p = ORDER > 0 ? ORDER : it.order, nf = ORDER > 0 ? 2 * ORDER - 1 : it.nf;
  int P = 1, F = p;
  for (int d = 0; d < DIM; ++d) P *= p;
  for (int d = 0; d + 1 < DIM; ++d) F *= nf;
  double2* s_tw = sm2;
  double2* bufA = gscratch ? gscratch + static_cast<size_t>(blockIdx.x) * 2 * scratch_elems : sm2 + nf;
  double2* bufB = bufA + scratch_elems;
  for (int i = threadIdx.x; i < nf; i += kBlock) s_tw[i] = it.tw[i];
  __syncthreads();
  const int total = a.n_active * NC * a.kn;
  for (int w = blockIdx.x; w < total; w += gridDim.x) {
    const int b = w % a.kn;
    const int ct = (w / a.kn) % NC;
    const int slot = w / (a.kn * NC);
    if (!((a.trg_mask[slot] >> ct) & 1u)) continue;  // uniform across the CTA
    const double2* in0 = a.Lhat + ((static_cast<size_t>(slot) * NC + ct) * a.kn + b) * F;
    double* Lc;
    if (a.L) {
      const int pidx = a.active[slot], lvl = m2l_slot_level(a, slot);
      const uint32_t pkey = a.trg.keys[a.trg.cell_off[lvl - 1] + pidx];
      const int cidx = a.trg.dense[a.trg.dense_off[lvl] + ((pkey << DIM) | ct)];
      Lc = a.L + (static_cast<size_t>(a.trg.cell_off[lvl] + cidx) * a.kn + b) * P;
    } else {
      Lc = a.Lc + ((static_cast<size_t>(slot) * NC + ct) * a.kn + b) * P;
    }
    if constexpr (DIM == 1) {
      idft_stage_c2r(in0, Lc, 1, p, s_tw, nf, false);
    } else {
      const double2* in = in0;
      double2* ob = bufA;
      int outer = 1;
      int inner = F / nf;  // nf^(DIM-2) * p
      for (int d = 0; d + 1 < DIM; ++d) {
        dft_stage<false>(in, ob, outer, nf, p, inner, s_tw, nf, true);
        __syncthreads();
        in = ob;
        ob = ob == bufA ? bufB : bufA;
        outer *= p;
        inner /= nf;
      }
      idft_stage_c2r(in, Lc, P / p, p, s_tw, nf, false);
    }
    __syncthreads();
  }
}

// Register-blocked 3-D inverse DFT (orders compiled in).  Every thread owns one *column* of a
// stage: it loads the column's nf (or p) inputs once into registers and produces all p outputs,
// with the twiddles as kernel-parameter constants (constant-bank operands of the DFMAs), so the
// shared-memory traffic is one read and one write per element and stage instead of two reads
// per complex multiply-add.  A CTA transforms NB cells at a time.
struct TwTable {
  double2 w[2 * 12 - 1];  // forward twiddles (cos, -sin) of length nf = 2 * order - 1, order <= 12
};

template <int ORDER, int NB>
__global__ void __launch_bounds__(256) k_m2l_idft3(M2LArgs a, TwTable tw) {
  constexpr int DIM = 3, NC = 8, p = ORDER, nf = 2 * ORDER - 1;
  constexpr int P = p * p * p, F = nf * nf * p, YN = p * nf * p;
  extern __shared__ double2 sm2[];
  double2* Y = sm2;            // [NB][p][nf][p]
  double2* Z = Y + NB * YN;    // [NB][p][p][p]
  __shared__ const double2* s_in[NB];
  __shared__ double* s_out[NB];
  const int total = a.n_active * NC * a.kn;
  const int w0 = blockIdx.x * NB;
  if (threadIdx.x < NB) {
    const int w = w0 + threadIdx.x;
    const double2* in = nullptr;
    double* out = nullptr;
    if (w < total) {
      const int b = w % a.kn;
      const int ct = (w / a.kn) % NC;
      const int slot = w / (a.kn * NC);
      if ((a.trg_mask[slot] >> ct) & 1u) {
        in = a.Lhat + ((static_cast<size_t>(slot) * NC + ct) * a.kn + b) * F;
        if (a.L) {
          const int pidx = a.active[slot], lvl = m2l_slot_level(a, slot);
          const uint32_t pkey = a.trg.keys[a.trg.cell_off[lvl - 1] + pidx];
          const int cidx = a.trg.dense[a.trg.dense_off[lvl] + ((pkey << DIM) | ct)];
          out = a.L + (static_cast<size_t>(a.trg.cell_off[lvl] + cidx) * a.kn + b) * P;
        } else {
          out = a.Lc + ((static_cast<size_t>(slot) * NC + ct) * a.kn + b) * P;
        }
      }
    }
    s_in[threadIdx.x] = in;
    s_out[threadIdx.x] = out;
  }
  __syncthreads();
  // stage A: axis 0, nf -> p.  column = (f1, f2)
  for (int item = threadIdx.x; item < NB * nf * p; item += blockDim.x) {
    const int c = item / (nf * p), col = item % (nf * p);
    const double2* in = s_in[c];
    if (!in) continue;
    double2 x[nf];
#pragma unroll
    for (int f = 0; f < nf; ++f) {
      // read once.  The streaming hint pays from order 8 on (order 12: 2.40 -> 1.67 ms); at order 6 the tail of the
      // spectra the Hadamard kernel has just written is still in L2 and a plain load is faster (1.59 vs 1.70 ms).
      const double2* q = in + f * (nf * p) + col;
      x[f] = ORDER >= 8 ? __ldcs(q) : *q;
    }
#pragma unroll
    for (int m = 0; m < p; ++m) {
      double re = 0.0, im = 0.0;
#pragma unroll
      for (int f = 0; f < nf; ++f) {
        const double2 t = tw.w[(m * f) % nf];  // conj: e^{+i theta}
        re = fma(x[f].x, t.x, re);
        re = fma(x[f].y, t.y, re);
        im = fma(x[f].y, t.x, im);
        im = fma(-x[f].x, t.y, im);
      }
      Y[c * YN + m * (nf * p) + col] = make_double2(re, im);
    }
  }
  __syncthreads();
  // stage B: axis 1, nf -> p.  column = (m0, f2)
  for (int item = threadIdx.x; item < NB * p * p; item += blockDim.x) {
    const int c = item / (p * p), col = item % (p * p);
    if (!s_in[c]) continue;
    const int m0 = col / p, f2 = col % p;
    const double2* yin = Y + c * YN + m0 * (nf * p) + f2;
    double2 x[nf];
#pragma unroll
    for (int f = 0; f < nf; ++f) x[f] = yin[f * p];
#pragma unroll
    for (int m = 0; m < p; ++m) {
      double re = 0.0, im = 0.0;
#pragma unroll
      for (int f = 0; f < nf; ++f) {
        const double2 t = tw.w[(m * f) % nf];
        re = fma(x[f].x, t.x, re);
        re = fma(x[f].y, t.y, re);
        im = fma(x[f].y, t.x, im);
        im = fma(-x[f].x, t.y, im);
      }
      Z[c * P + (m0 * p + m) * p + f2] = make_double2(re, im);
    }
  }
  __syncthreads();
  // stage C: axis 2, half spectrum -> real.  column = (m0, m1)
  for (int item = threadIdx.x; item < NB * p * p; item += blockDim.x) {
    const int c = item / (p * p), col = item % (p * p);
    double* out = s_out[c];
    if (!out) continue;
    const double2* zin = Z + c * P + col * p;
    double2 x[p];
#pragma unroll
    for (int k = 0; k < p; ++k) x[k] = zin[k];
#pragma unroll
    for (int m = 0; m < p; ++m) {
      double acc = x[0].x;
#pragma unroll
      for (int k = 1; k < p; ++k) {
        const double2 t = tw.w[(k * m) % nf];
        acc = fma(2.0 * x[k].x, t.x, acc);
        acc = fma(2.0 * x[k].y, t.y, acc);
      }
      out[col * p + m] = acc;
    }
  }
}

// Register-blocked 3-D forward DFT of the multipoles (the mirror image of k_m2l_idft3):
//   stage A: axis 2, p real -> p complex (half spectrum);  B: axis 1, p -> nf;  C: axis 0, p -> nf.
// Zero padding from p to nf points is implicit (only p inputs per column are read).
template <int ORDER, int NB, bool FLAGGED>
__global__ void __launch_bounds__(256) k_m2hat3(int first_cell, int n_cells, int km, const double* __restrict__ M,
                                                double2* __restrict__ Mhat, TwTable tw,
                                                const unsigned char* __restrict__ flags) {
  constexpr int p = ORDER, nf = 2 * ORDER - 1;
  constexpr int P = p * p * p, F = nf * nf * p, YN = p * nf * p;
  extern __shared__ double2 sm2[];
  double2* Y1 = sm2;            // [NB][p][p][p]   (n0, n1, k2)
  double2* Y2 = Y1 + NB * P;    // [NB][p][nf][p]  (n0, k1, k2)
  const int total = n_cells * km;
  const int w0 = blockIdx.x * NB;
  const int nb = min(NB, total - w0);
  // partitioned upward pass: skip the cells whose spectrum this rank does not need
  __shared__ int s_on[NB];
  if constexpr (FLAGGED) {
    if (threadIdx.x < NB)
      s_on[threadIdx.x] = threadIdx.x < nb && (flags[first_cell + (w0 + threadIdx.x) / km] & kCellFlagMhat);
    __syncthreads();
    bool any = false;
    for (int c = 0; c < nb; ++c) any = any || s_on[c];
    if (!any) return;
  }
  const double* Mc = M + static_cast<size_t>(first_cell) * km * P + static_cast<size_t>(w0) * P;
  double2* out = Mhat + static_cast<size_t>(w0) * F;
  // stage A: column = (n0, n1)
  for (int item = threadIdx.x; item < nb * p * p; item += blockDim.x) {
    const int c = item / (p * p), col = item % (p * p);
    if (FLAGGED && !s_on[c]) continue;
    const double* in = Mc + static_cast<size_t>(c) * P + col * p;
    double x[p];
#pragma unroll
    for (int n = 0; n < p; ++n) x[n] = in[n];
#pragma unroll
    for (int k = 0; k < p; ++k) {
      double re = 0.0, im = 0.0;
#pragma unroll
      for (int n = 0; n < p; ++n) {
        const double2 t = tw.w[(k * n) % nf];
        re = fma(x[n], t.x, re);
        im = fma(x[n], t.y, im);
      }
      Y1[c * P + col * p + k] = make_double2(re, im);
    }
  }
  __syncthreads();
  // stage B: column = (n0, k2), stride p
  for (int item = threadIdx.x; item < nb * p * p; item += blockDim.x) {
    const int c = item / (p * p), col = item % (p * p);
    if (FLAGGED && !s_on[c]) continue;
    const int n0 = col / p, k2 = col % p;
    const double2* in = Y1 + c * P + n0 * p * p + k2;
    double2 x[p];
#pragma unroll
    for (int n = 0; n < p; ++n) x[n] = in[n * p];
#pragma unroll
    for (int k = 0; k < nf; ++k) {
      double re = 0.0, im = 0.0;
#pragma unroll
      for (int n = 0; n < p; ++n) {
        const double2 t = tw.w[(k * n) % nf];
        re = fma(x[n].x, t.x, re);
        re = fma(-x[n].y, t.y, re);
        im = fma(x[n].x, t.y, im);
        im = fma(x[n].y, t.x, im);
      }
      Y2[c * YN + (n0 * nf + k) * p + k2] = make_double2(re, im);
    }
  }
  __syncthreads();
  // stage C: column = (k1, k2), stride nf * p
  for (int item = threadIdx.x; item < nb * nf * p; item += blockDim.x) {
    const int c = item / (nf * p), col = item % (nf * p);
    if (FLAGGED && !s_on[c]) continue;
    const double2* in = Y2 + c * YN + col;
    double2 x[p];
#pragma unroll
    for (int n = 0; n < p; ++n) x[n] = in[n * (nf * p)];
    double2* o = out + static_cast<size_t>(c) * F + col;
#pragma unroll
    for (int k = 0; k < nf; ++k) {
      double re = 0.0, im = 0.0;
#pragma unroll
      for (int n = 0; n < p; ++n) {
        const double2 t = tw.w[(k * n) % nf];
        re = fma(x[n].x, t.x, re);
        re = fma(-x[n].y, t.y, re);
        im = fma(x[n].x, t.y, im);
        im = fma(x[n].y, t.x, im);
      }
      o[static_cast<size_t>(k) * (nf * p)] = make_double2(re, im);
    }
  }
}

// Work counters for the roofline figures (not on the timed path): M2L pairs and target cells
// with a non-empty list at one level; P2P pairs at the leaves.
template <int DIM>
__global__ void k_count_m2l(TreeView src, TreeView trg, int level, unsigned long long* __restrict__ counters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long pairs = 0;
  if (i < trg.n_cells[level]) {
    int c[DIM], pc[DIM];
    morton_decode<DIM>(trg.keys[trg.cell_off[level] + i], c);
    const int nside_p = 1 << (level - 1);
    const int* sd = src.dense + src.dense_off[level];
    int nn = 1;
    for (int a = 0; a < DIM; ++a) { nn *= 3; pc[a] = c[a] >> 1; }
    for (int e = 0; e < nn; ++e) {
      int q[DIM], r = e;
      bool ok = true;
#pragma unroll
      for (int a = DIM - 1; a >= 0; --a) {
        q[a] = pc[a] + (r % 3) - 1;
        r /= 3;
        ok = ok && q[a] >= 0 && q[a] < nside_p;
      }
      if (!ok) continue;
      for (int ch = 0; ch < (1 << DIM); ++ch) {
        int s[DIM];
        bool far = false;
#pragma unroll
        for (int a = 0; a < DIM; ++a) {
          s[a] = 2 * q[a] + ((ch >> (DIM - 1 - a)) & 1);
          int o = s[a] - c[a];
          far = far || o > 1 || o < -1;
        }
        if (far && sd[morton_encode<DIM>(s)] >= 0) ++pairs;
      }
    }
  }
  if (pairs) {
    atomicAdd(&counters[0], pairs);
    atomicAdd(&counters[1], 1ull);
  }
}

template <int DIM>
__global__ void k_count_p2p(TreeView src, TreeView trg, unsigned long long* __restrict__ counters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int leaf = trg.height - 1;
  if (i >= trg.n_cells[leaf]) return;
  int c[DIM];
  morton_decode<DIM>(trg.keys[trg.cell_off[leaf] + i], c);
  const int nside = 1 << leaf;
  const int* sd = src.dense + src.dense_off[leaf];
  int nn = 1;
  for (int a = 0; a < DIM; ++a) nn *= 3;
  unsigned long long ns = 0;
  for (int e = 0; e < nn; ++e) {
    int q[DIM], r = e;
    bool ok = true;
#pragma unroll
    for (int a = DIM - 1; a >= 0; --a) {
      q[a] = c[a] + (r % 3) - 1;
      r /= 3;
      ok = ok && q[a] >= 0 && q[a] < nside;
    }
    if (!ok) continue;
    int sc = sd[morton_encode<DIM>(q)];
    if (sc >= 0) ns += src.leaf_start[sc + 1] - src.leaf_start[sc];
  }
  unsigned long long nt = trg.leaf_start[i + 1] - trg.leaf_start[i];
  if (ns) atomicAdd(&counters[2], ns * nt);
}

size_t smem_opt_in(const void* fn, size_t bytes) {
  if (bytes > 40 * 1024) {  // static shared memory counts towards the 48 KiB default limit
    PLT_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
  }
  return bytes;
}

constexpr size_t kSmemCap = 200 * 1024;

struct InterpTablesView {
  const double* tw;  // host copy of the forward twiddles [nf][2]
};

// (dim, order) -> template instance.  The orders of the two fixed policies of the reference
// (accuracy = infinity -> 6, accuracy = 0 -> 12, src/fmm/fmm_accuracy_estimator.hpp:76-82) and the
// first steps of the search are compiled with the order as a constant (index arithmetic without
// integer divisions, unrolled contractions); any other order runs the generic instance (ORDER = 0).
template <class F>
void dispatch_dim_order(int dim, int order, F&& f) {
  auto with_order = [&](auto dm) {
    if constexpr (decltype(dm)::value == 1) {
      f(dm, std::integral_constant<int, 0>{});
    } else {
      switch (order) {
        case 6: f(dm, std::integral_constant<int, 6>{}); break;
        case 8: f(dm, std::integral_constant<int, 8>{}); break;
        case 10: f(dm, std::integral_constant<int, 10>{}); break;
        case 12: f(dm, std::integral_constant<int, 12>{}); break;
        default: f(dm, std::integral_constant<int, 0>{}); break;
      }
    }
  };
  switch (dim) {
    case 1: with_order(std::integral_constant<int, 1>{}); break;
    case 2: with_order(std::integral_constant<int, 2>{}); break;
    case 3: with_order(std::integral_constant<int, 3>{}); break;
    default: throw Error(PLT_ERR_INVALID, "dim must be 1, 2 or 3");
  }
}

}  // namespace

// ------------------------------------------------------------------------------------
// Launchers
// ------------------------------------------------------------------------------------
void launch_transform_points(int dim, const double* aniso, const double* pts, int64_t n, double* pos,
                             cudaStream_t s, LaunchCounter& c, int64_t ld) {
  if (n == 0) return;
  PLT_LAUNCH(c, k_transform_points, ceil_div(n, 256), 256, 0, s, dim, make_mat3(dim, aniso), pts, n, ld > 0 ? ld : n,
             pos);
}

void launch_prepare_weights(int kind, int dim, const double* aniso, const double* w, const int* perm, int64_t n,
                            double* wt, cudaStream_t s, LaunchCounter& c) {
  if (n == 0) return;
  const int fold = kind == KIND_F || kind == KIND_H;
  const int km = fold ? dim : 1;
  PLT_LAUNCH(c, k_prepare_weights, ceil_div(n, 256), 256, 0, s, km, fold, dim, make_mat3(dim, aniso), w, perm, n, wt);
}

void launch_finish_outputs(int kind, int dim, const double* aniso, const double* vt, const int* perm, int64_t n,
                           int64_t lo, int64_t hi, double* out, cudaStream_t s, LaunchCounter& c) {
  if (n == 0) return;
  const int fold = kind == KIND_FT || kind == KIND_H;
  const int kn = fold ? dim : 1;
  PLT_LAUNCH(c, k_finish_outputs, ceil_div(n, 256), 256, 0, s, kn, fold, dim, make_mat3(dim, aniso), vt, perm, n,
             lo, hi, out);
}

void launch_p2m(int dim, int km, const TreeView& tr, const Box& box, const InterpDev& it, const double* wt,
                double* M, cudaStream_t s, LaunchCounter& c) {
  const int leaf = tr.height - 1;
  const int n = tr.n_cells[leaf];
  if (n == 0) return;
  size_t smem = sizeof(double) * (it.order + kPointBatch * dim * it.order + kPointBatch * km);
  dispatch_dim_order(dim, it.order, [&](auto dm, auto od) {
    PLT_LAUNCH(c, (k_p2m<dm.value, od.value>), n, kBlock, smem, s, tr, box, it, km, wt, M);
  });
}

void launch_m2m(int dim, int km, const TreeView& tr, int level, const InterpDev& it, double* M, cudaStream_t s,
                LaunchCounter& c) {
  const int n = tr.n_cells[level];
  if (n == 0) return;
  const int P = nodes_per_cell(it.order, dim);
  size_t smem = sizeof(double) * (2 * it.order * it.order + 3 * static_cast<size_t>(P));
  PLT_REQUIRE(smem <= kSmemCap, "interpolation order too large for M2M shared-memory staging");
  dispatch_dim_order(dim, it.order, [&](auto dm, auto od) {
    smem_opt_in((const void*)k_m2m<dm.value, od.value>, smem);
    PLT_LAUNCH(c, (k_m2m<dm.value, od.value>), n, kBlock, smem, s, tr, level, it, km, M);
  });
}

void launch_l2l(int dim, int kn, const TreeView& tr, int level, const InterpDev& it, double* L, int cell_lo,
                int cell_hi, int par_lo, int par_hi, cudaStream_t s, LaunchCounter& c) {
  const int n = cell_hi - cell_lo;
  if (n <= 0) return;
  static const bool no_parent = getenv("PLT_DEBUG_NO_L2L_PARENT") != nullptr;  // A/B switch
  const int p = it.order;
  if (dim == 3 && !no_parent && it.host_child && par_hi > par_lo && (p == 6 || p == 8 || p == 10 || p == 12)) {
    LeafTables tb{};
    for (int i = 0; i < 2 * p * p; ++i) tb.child[i] = it.host_child[i];
    auto run = [&](auto od) {
      constexpr int P_ = od.value;
      const size_t bytes = sizeof(double) * 7 * P_ * P_ * P_;
      smem_opt_in((const void*)k_l2l_parent3<P_>, bytes);
      PLT_LAUNCH(c, (k_l2l_parent3<P_>), par_hi - par_lo, leaf_threads(P_), bytes, s, tr, level, tb, kn, L, par_lo,
                 cell_lo, cell_hi);
    };
    if (p == 6) run(std::integral_constant<int, 6>{});
    if (p == 8) run(std::integral_constant<int, 8>{});
    if (p == 10) run(std::integral_constant<int, 10>{});
    if (p == 12) run(std::integral_constant<int, 12>{});
    return;
  }
  const int P = nodes_per_cell(it.order, dim);
  size_t smem = sizeof(double) * (2 * it.order * it.order + 2 * static_cast<size_t>(P));
  PLT_REQUIRE(smem <= kSmemCap, "interpolation order too large for L2L shared-memory staging");
  dispatch_dim_order(dim, it.order, [&](auto dm, auto od) {
    smem_opt_in((const void*)k_l2l<dm.value, od.value>, smem);
    PLT_LAUNCH(c, (k_l2l<dm.value, od.value>), n, kBlock, smem, s, tr, level, it, kn, L, cell_lo);
  });
}

void launch_l2p(int dim, int kn, const TreeView& tr, const Box& box, const InterpDev& it, const double* L,
                double* vt, int64_t leaf_lo, int64_t leaf_hi, cudaStream_t s, LaunchCounter& c) {
  const int n = static_cast<int>(leaf_hi - leaf_lo);
  if (n <= 0) return;
  size_t smem = sizeof(double) * (it.order + (kBlock / 32) * dim * it.order);
  if (dim == 1) PLT_LAUNCH(c, k_l2p<1>, n, kBlock, smem, s, tr, box, it, kn, L, vt, static_cast<int>(leaf_lo));
  if (dim == 2) PLT_LAUNCH(c, k_l2p<2>, n, kBlock, smem, s, tr, box, it, kn, L, vt, static_cast<int>(leaf_lo));
  if (dim == 3) PLT_LAUNCH(c, k_l2p<3>, n, kBlock, smem, s, tr, box, it, kn, L, vt, static_cast<int>(leaf_lo));
}

namespace {
template <class F>
void dispatch_leaf(int dim, int order, F&& f) {
  auto with_order = [&](auto dm) {
    switch (order) {
      case 6: f(dm, std::integral_constant<int, 6>{}); break;
      case 8: f(dm, std::integral_constant<int, 8>{}); break;
      case 10: f(dm, std::integral_constant<int, 10>{}); break;
      case 12: f(dm, std::integral_constant<int, 12>{}); break;
      default: throw Error(PLT_ERR_INVALID, "fused leaf pass: order not compiled in");
    }
  };
  switch (dim) {
    case 1: with_order(std::integral_constant<int, 1>{}); break;
    case 2: with_order(std::integral_constant<int, 2>{}); break;
    case 3: with_order(std::integral_constant<int, 3>{}); break;
    default: throw Error(PLT_ERR_INVALID, "dim must be 1, 2 or 3");
  }
}
}  // namespace

bool launch_l2l_l2p_leaf(int dim, int kn, const TreeView& tr, const Box& box, const InterpDev& it, const double* L,
                         const double* Lc, const int* leaf_meta, double* vt, int64_t leaf_lo, int64_t leaf_hi,
                         int par_lo, int par_hi, cudaStream_t s, LaunchCounter& c) {
  const int n = par_hi - par_lo;
  if (n <= 0) return true;
  if (!leaf_fused_supported(dim, it.order) || !it.host_child || !it.host_beta) return false;
  const size_t smem = leaf_fused_smem_bytes(dim, it.order);
  LeafTables tb{};
  const int p = it.order;
  for (int i = 0; i < 2 * p * p; ++i) tb.child[i] = it.host_child[i];
  for (int i = 0; i < p; ++i) tb.beta[i] = it.host_beta[i];
  const int lo = static_cast<int>(leaf_lo), hi = static_cast<int>(leaf_hi);
  static const bool no_direct = getenv("PLT_DEBUG_NO_LEAF_DIRECT") != nullptr;  // A/B switch
  if (dim == 3 && it.polynomial && !no_direct && (p == 6 || p == 8 || p == 10)) {
    auto run = [&](auto od) {
      constexpr int P_ = od.value;
      const size_t bytes = sizeof(double) * (9 * P_ * (P_ * P_ + 2) + kLeafDirectCap * 2 * 3 * P_);
      smem_opt_in((const void*)k_leaf_direct3<P_>, bytes);
      PLT_LAUNCH(c, (k_leaf_direct3<P_>), n, kLeafDirectThreads, bytes, s, tr, box, tb, kn, L, Lc, leaf_meta, vt,
                 par_lo, lo, hi);
    };
    if (p == 6) run(std::integral_constant<int, 6>{});
    if (p == 8) run(std::integral_constant<int, 8>{});
    if (p == 10) run(std::integral_constant<int, 10>{});
    return true;
  }
  dispatch_leaf(dim, it.order, [&](auto dm, auto od) {
    smem_opt_in((const void*)k_l2l_l2p_leaf<dm.value, od.value>, smem);
    PLT_LAUNCH(c, (k_l2l_l2p_leaf<dm.value, od.value>), n, leaf_threads(od.value), smem, s, tr, box, tb, kn, L, Lc, leaf_meta, vt,
               par_lo, lo, hi);
  });
  return true;
}

size_t leaf_fused_smem_bytes(int dim, int order) {
  const size_t P = nodes_per_cell(order, dim);
  const size_t nc = size_t{1} << dim;
  const size_t PS = dim == 3 ? static_cast<size_t>(order) * (order * order + 1) : P;  // padded child stride
  const size_t last = dim == 1 ? 1 : (dim == 2 ? 2 : 4);
  return sizeof(double) * (nc * PS + last * P);
}

bool leaf_fused_supported(int dim, int order) {
  const bool compiled = order == 6 || order == 8 || order == 10 || order == 12;
  return compiled && leaf_fused_smem_bytes(dim, order) <= kSmemCap;
}

namespace {
// Scratch policy for the DFT kernels: shared memory when the two complex stage buffers fit,
// otherwise a per-CTA slice of a global scratch buffer (generic addressing, same code).
struct DftScratch {
  int elems;           // complex elements per stage buffer
  size_t smem;         // dynamic shared memory bytes
  bool global;
  int grid;
  DevBuf<double2> buf;
};

DftScratch plan_dft_scratch(int order, int dim, int work_items, cudaStream_t s) {
  DftScratch d;
  const int nf = 2 * order - 1;
  // largest intermediate: order^(dim-1) * nf^(dim-2)... bounded by order * nf^(dim-2) * order * ... use generous bound
  int elems = order;
  for (int a = 0; a + 1 < dim; ++a) elems *= (a == 0 ? order : nf);
  // dim 1: order; dim 2: order*order; dim 3: order*order*nf
  d.elems = elems;
  size_t smem = sizeof(double2) * (nf + 2 * static_cast<size_t>(elems));
  d.global = smem > kSmemCap;
  d.smem = d.global ? sizeof(double2) * nf : smem;
  d.grid = std::max(1, std::min(work_items, d.global ? 4 * num_sm() : work_items));
  if (d.global) d.buf.alloc(static_cast<size_t>(d.grid) * 2 * elems, s);
  return d;
}
}  // namespace

namespace {
template <int ORDER, int NB>
void launch_m2hat3(int first, int n_cells, int km, const double* tw_host, const double* M, double2* Mhat,
                   const unsigned char* flags, cudaStream_t s, LaunchCounter& c) {
  constexpr int p = ORDER, nf = 2 * ORDER - 1;
  TwTable tw{};
  for (int i = 0; i < nf; ++i) tw.w[i] = make_double2(tw_host[2 * i], tw_host[2 * i + 1]);
  const size_t smem = sizeof(double2) * NB * (p * p * p + p * nf * p);
  if (flags) {
    smem_opt_in((const void*)k_m2hat3<ORDER, NB, true>, smem);
    PLT_LAUNCH(c, (k_m2hat3<ORDER, NB, true>), ceil_div(n_cells * km, NB), 256, smem, s, first, n_cells, km, M, Mhat, tw,
               flags);
  } else {
    smem_opt_in((const void*)k_m2hat3<ORDER, NB, false>, smem);
    PLT_LAUNCH(c, (k_m2hat3<ORDER, NB, false>), ceil_div(n_cells * km, NB), 256, smem, s, first, n_cells, km, M, Mhat, tw,
               flags);
  }
}
}  // namespace

void launch_m2hat(int dim, int km, const TreeView& tr, const InterpDev& it, const double* M, double2* Mhat,
                  cudaStream_t s, LaunchCounter& c) {
  if (tr.height <= 2) return;
  const int first = tr.cell_off[2];
  int n_cells = 0;
  for (int l = 2; l < tr.height; ++l) n_cells += tr.n_cells[l];
  if (n_cells == 0) return;
  static const bool no_reg = getenv("PLT_DEBUG_NO_REGDFT") != nullptr;  // A/B switch for parity bisection
  if (dim == 3 && it.host_tw && !no_reg) {
    switch (it.order) {
      case 6: launch_m2hat3<6, 4>(first, n_cells, km, it.host_tw, M, Mhat, tr.flags, s, c); return;
      case 8: launch_m2hat3<8, 2>(first, n_cells, km, it.host_tw, M, Mhat, tr.flags, s, c); return;
      case 10: launch_m2hat3<10, 1>(first, n_cells, km, it.host_tw, M, Mhat, tr.flags, s, c); return;
      case 12: launch_m2hat3<12, 1>(first, n_cells, km, it.host_tw, M, Mhat, tr.flags, s, c); return;
      default: break;
    }
  }
  DftScratch d = plan_dft_scratch(it.order, dim, n_cells * km, s);
  dispatch_dim_order(dim, it.order, [&](auto dm, auto od) {
    smem_opt_in((const void*)k_m2hat<dm.value, od.value>, d.smem);
    PLT_LAUNCH(c, (k_m2hat<dm.value, od.value>), d.grid, kBlock, d.smem, s, first, n_cells, it, km, M, Mhat, d.buf.get(),
               d.elems, tr.flags);
  });
}

void launch_count_work(int dim, const TreeView& src, const TreeView& trg, unsigned long long* counters,
                       cudaStream_t s, LaunchCounter& c) {
  for (int l = 2; l < trg.height; ++l) {
    const int n = trg.n_cells[l];
    if (n == 0) continue;
    if (dim == 1) PLT_LAUNCH(c, k_count_m2l<1>, ceil_div(n, 256), 256, 0, s, src, trg, l, counters);
    if (dim == 2) PLT_LAUNCH(c, k_count_m2l<2>, ceil_div(n, 256), 256, 0, s, src, trg, l, counters);
    if (dim == 3) PLT_LAUNCH(c, k_count_m2l<3>, ceil_div(n, 256), 256, 0, s, src, trg, l, counters);
  }
  const int n = trg.n_cells[trg.height - 1];
  if (n == 0) return;
  if (dim == 1) PLT_LAUNCH(c, k_count_p2p<1>, ceil_div(n, 256), 256, 0, s, src, trg, counters);
  if (dim == 2) PLT_LAUNCH(c, k_count_p2p<2>, ceil_div(n, 256), 256, 0, s, src, trg, counters);
  if (dim == 3) PLT_LAUNCH(c, k_count_p2p<3>, ceil_div(n, 256), 256, 0, s, src, trg, counters);
}

namespace {
template <int DIM>
void launch_hadamard_tiled(const M2LArgs& a, int F, cudaStream_t s, LaunchCounter& c) {
  constexpr int NN = M2LGeom<DIM>::NN, NC = M2LGeom<DIM>::NC, NOFF = M2LGeom<DIM>::NOFF;
  constexpr int kHadWarps = HadCfg<DIM>::kWarps;
  const int n_ftiles = ceil_div(F, kHadTF);
  // one persistent CTA per SM; fewer when there is not a warp-round of parents per CTA
  const long long rounds = static_cast<long long>(n_ftiles) * ceil_div(a.n_active, kHadWarps);
  const int grid = static_cast<int>(std::max<long long>(1, std::min<long long>(num_sm(), rounds)));
  const size_t smem = sizeof(double2) * NOFF * kHadTF + sizeof(int2) * (NN * NC + (NN * NC - NC) * kHadWarps);
  // rows of a tile start at multiples of 512 B of a 16 F-byte row: the bulk copies need 16-byte alignment (always) --
  // PLT_DEBUG_NO_TMA keeps the register-staged loop for A/B runs
  static const int no_tma = getenv("PLT_DEBUG_NO_TMA") != nullptr ? 1 : 0;
  if (a.kn * a.km == 1) {
    smem_opt_in((const void*)k_m2l_hadamard_tiled<DIM, false>, smem);
    PLT_LAUNCH(c, (k_m2l_hadamard_tiled<DIM, false>), grid, kHadWarps * 32, smem, s, a, F, n_ftiles, no_tma);
  } else {
    smem_opt_in((const void*)k_m2l_hadamard_tiled<DIM, true>, smem);
    PLT_LAUNCH(c, (k_m2l_hadamard_tiled<DIM, true>), grid, kHadWarps * 32, smem, s, a, F, n_ftiles, no_tma);
  }
}
}  // namespace

bool m2l_hadamard_multi_level_supported() {
  static const bool ok = getenv("PLT_DEBUG_NO_TILED") == nullptr && getenv("PLT_DEBUG_NO_MULTI_LEVEL") == nullptr;
  return ok && !hadamard_tmem_enabled();
}

void launch_m2l_hadamard(const M2LArgs& a, cudaStream_t s, LaunchCounter& c) {
  if (a.n_active == 0) return;
  PLT_REQUIRE(a.n_lvls == 1 || m2l_hadamard_multi_level_supported(), "multi-level Hadamard launch not available");
  const int F = freqs_per_cell(a.order, a.dim);
  static const bool no_tiled = getenv("PLT_DEBUG_NO_TILED") != nullptr;  // A/B switch for parity bisection
  if (!no_tiled) {
    if (launch_m2l_hadamard_tmem(a, F, s, c)) return;
    if (a.dim == 1) launch_hadamard_tiled<1>(a, F, s, c);
    if (a.dim == 2) launch_hadamard_tiled<2>(a, F, s, c);
    if (a.dim == 3) launch_hadamard_tiled<3>(a, F, s, c);
    return;
  }
  const int threads = F >= 256 ? 256 : ((F + 31) / 32 * 32);
#define PLT_HAD(D, KN, KM) PLT_LAUNCH(c, (k_m2l_hadamard<D, KN, KM>), a.n_active, threads, 0, s, a, F)
  const int key = a.dim * 100 + a.kn * 10 + a.km;
  switch (key) {
    case 111: PLT_HAD(1, 1, 1); break;
    case 211: PLT_HAD(2, 1, 1); break;
    case 311: PLT_HAD(3, 1, 1); break;
    case 212: PLT_HAD(2, 1, 2); break;
    case 221: PLT_HAD(2, 2, 1); break;
    case 222: PLT_HAD(2, 2, 2); break;
    case 313: PLT_HAD(3, 1, 3); break;
    case 331: PLT_HAD(3, 3, 1); break;
    case 333: PLT_HAD(3, 3, 3); break;
    default: throw Error(PLT_ERR_INVALID, "unsupported (dim, kn, km)");
  }
#undef PLT_HAD
}

namespace {
template <int ORDER, int NB>
void launch_idft3(const M2LArgs& a, const InterpTablesView& tv, cudaStream_t s, LaunchCounter& c) {
  constexpr int p = ORDER, nf = 2 * ORDER - 1;
  TwTable tw{};
  for (int i = 0; i < nf; ++i) tw.w[i] = make_double2(tv.tw[2 * i], tv.tw[2 * i + 1]);
  const size_t smem = sizeof(double2) * NB * (p * nf * p + p * p * p);
  smem_opt_in((const void*)k_m2l_idft3<ORDER, NB>, smem);
  const int total = a.n_active * 8 * a.kn;
  PLT_LAUNCH(c, (k_m2l_idft3<ORDER, NB>), ceil_div(total, NB), 256, smem, s, a, tw);
}
}  // namespace

void launch_m2l_idft(const M2LArgs& a, const InterpDev& it, cudaStream_t s, LaunchCounter& c) {
  if (a.n_active == 0) return;
  static const bool no_reg = getenv("PLT_DEBUG_NO_REGDFT") != nullptr;  // A/B switch for parity bisection
  if (a.dim == 3 && it.host_tw && !no_reg) {
    const InterpTablesView tv{it.host_tw};
    switch (it.order) {
      case 6: launch_idft3<6, 4>(a, tv, s, c); return;
      case 8: launch_idft3<8, 2>(a, tv, s, c); return;
      case 10: launch_idft3<10, 1>(a, tv, s, c); return;
      case 12: launch_idft3<12, 1>(a, tv, s, c); return;
      default: break;
    }
  }
  const int work = a.n_active * (1 << a.dim) * a.kn;
  DftScratch d = plan_dft_scratch(it.order, a.dim, work, s);
  dispatch_dim_order(a.dim, it.order, [&](auto dm, auto od) {
    smem_opt_in((const void*)k_m2l_idft<dm.value, od.value>, d.smem);
    PLT_LAUNCH(c, (k_m2l_idft<dm.value, od.value>), d.grid, kBlock, d.smem, s, a, it, d.buf.get(), d.elems);
  });
}

}  // namespace plt
