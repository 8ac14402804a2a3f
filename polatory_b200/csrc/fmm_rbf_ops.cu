// RBF-dependent FMM kernels: leaf-level near field (P2P) and tabulation of the
// Fourier-space M2L operators.
#include "dispatch.cuh"
#include "fmm_ops.cuh"

namespace plt {
namespace {

// ------------------------------------------------------------------------------------
// P2P over the 3^dim adjacent source leaves (separation criterion 1,
// include/polatory/fmm/kernel.hpp:71; list = build_p2p_interaction_list(src, trg, 1, mutual),
// src/fmm/fmm_evaluator.hpp:95-97).
//
// One warp per *active* target leaf (plan.cuh: leaves with at least one non-empty adjacent
// source leaf).  The sources of all adjacent leaves are addressed as one concatenated list
// (27-entry prefix table per warp in shared memory, binary search per lane), so all 32 lanes
// own a source even when single leaves hold only a handful of points; loads of the
// Morton-sorted SoA arrays stay contiguous within a leaf.  Targets are processed in register
// groups of kTG; partial sums stay in registers over the whole source list and are reduced
// with warp shuffles once per group.  The symmetric variant needs no special casing: the
// i == j pair evaluated at d = 0 is exactly the k(0,0) w_i self term of
// src/fmm/fmm_symmetric_evaluator.hpp:163-193.
// ------------------------------------------------------------------------------------
constexpr int kP2PWarps = 4;
constexpr int kTG = 8;
#ifndef PLT_P2P_MINB_K
#define PLT_P2P_MINB_K 4
#endif
constexpr int kSrcCap = 384;  // sources staged per warp and chunk

// Value kernels whose pair loop is written out by hand (below): the slope and the smoothing constant are folded
// out of the loop (w' = -+s w per source; r^2 accumulated onto c^2 by the three FMAs), and the kTG / 2 independent
// pair evaluations of a half group are interleaved stage by stage.  The generic loop compiles to ONE dependent chain
// per pair at a time (19 instructions, 15 of them on the FP64 pipe, no ILP: profiles/r02_e_p2p.md), which left the
// FP64 pipe at 52 % with 21 resident warps.
template <int FAM, int KIND>
struct FoldedPair { static constexpr bool value = KIND == KIND_K && (FAM == FAM_BH3 || FAM == FAM_TH3); };

template <int FAM, int DIM, int N>
__device__ __forceinline__ void folded_pairs(double c1e, const double (&sp)[DIM], double w0,
                                             const double (*tp)[DIM], double (*v)[1]) {
  double r2[N], y[N], xy[N], e[N];
#pragma unroll
  for (int u = 0; u < N; ++u) {
    double acc = c1e;
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
      const double d = tp[u][a] - sp[a];
      acc = fma(d, d, acc);
    }
    r2[u] = acc;
  }
#pragma unroll
  for (int u = 0; u < N; ++u) y[u] = rsqrt_seed(r2[u]);
#pragma unroll
  for (int u = 0; u < N; ++u) xy[u] = r2[u] * y[u];
#pragma unroll
  for (int u = 0; u < N; ++u) e[u] = fma(-xy[u], y[u], 1.0);
#pragma unroll
  for (int u = 0; u < N; ++u) {
    const double q = fma(0.375, e[u], 0.5);
    const double rho = fma(xy[u] * e[u], q, xy[u]);  // sqrt(r^2 + c^2), as sqrt_fast
    if constexpr (FAM == FAM_BH3) {
      v[u][0] = fma(rho, w0, v[u][0]);               // phi = -s rho
    } else {
      v[u][0] = fma(r2[u] * rho, w0, v[u][0]);       // phi = +s rho^3
    }
  }
}

template <int FAM, int KIND, int DIM>
__global__ void __launch_bounds__(kP2PWarps * 32, KIND == KIND_K ? PLT_P2P_MINB_K : 4)
k_p2p(RbfConst k, TreeView src, const double* __restrict__ swt, TreeView trg, double* __restrict__ vt,
      const int* __restrict__ leaves, int n_leaves, int leaf_lo, int leaf_hi, int* __restrict__ queue,
      int n_split) {
  constexpr int KM = KindTraits<KIND, DIM>::km;
  constexpr int KN = KindTraits<KIND, DIM>::kn;
  constexpr int NN = DIM == 1 ? 3 : (DIM == 2 ? 9 : 27);
  __shared__ int s_start[kP2PWarps][32];   // first source point of neighbour nb
  __shared__ int s_prefix[kP2PWarps][32];  // exclusive prefix of the neighbour sizes; [NN] = total
  extern __shared__ double s_dat[];        // [warps][DIM + KM][kSrcCap] staged sources
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // Persistent warps pull target leaves from a global queue: leaf populations of surface clouds
  // vary by an order of magnitude, and a static leaf -> warp map left a third of the resident
  // warp slots idle (ncu: 15.6 of 24 warps active per SM).
  // A queue item is (leaf, split): the target groups of a heavy leaf are dealt round-robin to
  // n_split items so that a few hundred-point leaves (volume clouds under anisotropy) do not
  // serialise on one warp each.
  for (;;) {
  int qi = 0;
  if (lane == 0) qi = atomicAdd(queue, 1);
  qi = __shfl_sync(0xffffffffu, qi, 0);
  const int li = qi / n_split, split = qi - li * n_split;
  if (li >= n_leaves) break;
  const int cell = leaves[li];
  if (cell < leaf_lo || cell >= leaf_hi) continue;
  const int leaf = trg.height - 1;
  const int nside = 1 << leaf;
  int tc[DIM];
  morton_decode<DIM>(trg.keys[trg.cell_off[leaf] + cell], tc);
  const int t0 = trg.leaf_start[cell], t1 = trg.leaf_start[cell + 1];
  if (t0 + split * kTG >= t1) continue;
  const int* sdense = src.dense + src.dense_off[leaf];

  // neighbour table: lane nb < NN looks up its source leaf
  int s0 = 0, cnt = 0;
  if (lane < NN) {
    int q[DIM], r = lane;
    bool ok = true;
#pragma unroll
    for (int a = DIM - 1; a >= 0; --a) {
      q[a] = tc[a] + (r % 3) - 1;
      r /= 3;
      ok = ok && q[a] >= 0 && q[a] < nside;
    }
    if (ok) {
      const int sc = sdense[morton_encode<DIM>(q)];
      if (sc >= 0) {
        s0 = src.leaf_start[sc];
        cnt = src.leaf_start[sc + 1] - s0;
      }
    }
  }
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  __syncwarp();  // previous leaf done with the tables
  s_start[warp][lane] = s0;
  s_prefix[warp][lane] = incl - cnt;
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  __syncwarp();
  const int* pre = s_prefix[warp];
  const int* st = s_start[warp];

  // The concatenated source list is staged in shared memory once per (leaf, chunk of kSrcCap sources): SoA rows of
  // DIM coordinates + KM weights, lane-contiguous (conflict-free LDS.64).  Every target group of the leaf then
  // streams it from there: no prefix search and no global load inside the pair loop.
  double* dat = s_dat + static_cast<size_t>(warp) * (DIM + KM) * kSrcCap;
  // neighbour holding concatenated index jj (last nb with prefix[nb] <= jj) -> sorted source index
  auto source_index = [&](int jj) {
    int lo = 0, hi = NN - 1;
#pragma unroll
    for (int it = 0; it < 5; ++it) {
      const int mid = (lo + hi + 1) >> 1;
      if (pre[mid] <= jj) lo = mid; else hi = mid - 1;
    }
    return st[lo] + (jj - pre[lo]);
  };
  // the pair evaluations of one source against the target group in registers
  auto eval_source = [&](const double (&sp)[DIM], const double (&w)[KM], const double (*tp)[DIM], double (*v)[KN],
                         int nt) {
    if constexpr (FoldedPair<FAM, KIND>::value) {
      const double c1e = k.c[1] + 1e-300;  // c^2 (+ the guard that keeps the seed finite at r = 0)
      // (warp-uniform: slots up to the next multiple of 2 beyond nt are evaluated, not stored)
      if (nt > kTG / 2) {
        folded_pairs<FAM, DIM, kTG / 2>(c1e, sp, w[0], tp, v);
        if (nt > 3 * kTG / 4) folded_pairs<FAM, DIM, kTG / 2>(c1e, sp, w[0], tp + kTG / 2, v + kTG / 2);
        else folded_pairs<FAM, DIM, kTG / 4>(c1e, sp, w[0], tp + kTG / 2, v + kTG / 2);
      } else if (nt > kTG / 4) {
        folded_pairs<FAM, DIM, kTG / 2>(c1e, sp, w[0], tp, v);
      } else {
        folded_pairs<FAM, DIM, kTG / 4>(c1e, sp, w[0], tp, v);
      }
    } else if (nt == kTG) {
      // full group: branch-free, the kTG independent pair evaluations interleave
#pragma unroll
      for (int u = 0; u < kTG; ++u) {
        double d[DIM];
#pragma unroll
        for (int a = 0; a < DIM; ++a) d[a] = tp[u][a] - sp[a];
        pair_accumulate<FAM, KIND, DIM>(k, d, w, v[u]);
      }
    } else {
#pragma unroll
      for (int u = 0; u < kTG; ++u) {
        if (u < nt) {
          double d[DIM];
#pragma unroll
          for (int a = 0; a < DIM; ++a) d[a] = tp[u][a] - sp[a];
          pair_accumulate<FAM, KIND, DIM>(k, d, w, v[u]);
        }
      }
    }
  };
  auto load_targets = [&](int tb, double (*tp)[DIM], double (*v)[KN]) {
#pragma unroll
    for (int u = 0; u < kTG; ++u) {
      const int t = min(tb + u, t1 - 1);
#pragma unroll
      for (int a = 0; a < DIM; ++a) tp[u][a] = trg.pos[a * trg.n + t];
#pragma unroll
      for (int b = 0; b < KN; ++b) v[u][b] = 0.0;
    }
  };
  auto store_targets = [&](int tb, int nt, double (*v)[KN]) {
#pragma unroll
    for (int u = 0; u < kTG; ++u) {
#pragma unroll
      for (int b = 0; b < KN; ++b) {
        double x = v[u][b];
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0 && u < nt) vt[b * trg.n + tb + u] += x;
      }
    }
  };
  const double wscale = FoldedPair<FAM, KIND>::value ? (FAM == FAM_BH3 ? -k.c[0] : k.c[0]) : 1.0;  // slope folded in

  const int tb0 = t0 + split * kTG;
  if (tb0 + n_split * kTG >= t1) {
    // ONE target group for this queue item (a few targets per leaf: grids, coarse samplers): the sources are used
    // once, so they go from global memory straight into the pair loop.
    // (reduced and added per kSrcCap sources like the staged path below, so that a target gets the same bits whichever
    // path its leaf takes: per-layer batches of the isosurface sampler equal the one-shot evaluation exactly)
    const int nt = min(kTG, t1 - tb0);
    for (int c0 = 0; c0 < total; c0 += kSrcCap) {
      double tp[kTG][DIM], v[kTG][KN];
      load_targets(tb0, tp, v);
      const int c1 = min(total, c0 + kSrcCap);
      for (int jj = c0 + lane; jj < c1; jj += 32) {
        const int j = source_index(jj);
        double sp[DIM], w[KM];
#pragma unroll
        for (int a = 0; a < DIM; ++a) sp[a] = src.pos[a * src.n + j];
#pragma unroll
        for (int m = 0; m < KM; ++m) w[m] = swt[m * src.n + j] * wscale;
        eval_source(sp, w, tp, v, nt);
      }
      store_targets(tb0, nt, v);
    }
    continue;
  }
  // Several target groups: the concatenated source list is staged in shared memory once per (leaf, chunk of kSrcCap
  // sources): SoA rows of DIM coordinates + KM weights, lane-contiguous (conflict-free LDS.64).  Every target group
  // then streams it from there: no prefix search and no global load inside the pair loop.
  for (int c0 = 0; c0 < total; c0 += kSrcCap) {
    const int nc = min(kSrcCap, total - c0);
    __syncwarp();  // previous chunk no longer read
    for (int q = lane; q < nc; q += 32) {
      const int j = source_index(c0 + q);
#pragma unroll
      for (int a = 0; a < DIM; ++a) dat[a * kSrcCap + q] = src.pos[a * src.n + j];
#pragma unroll
      for (int m = 0; m < KM; ++m) dat[(DIM + m) * kSrcCap + q] = swt[m * src.n + j] * wscale;
    }
    __syncwarp();
    for (int tb = tb0; tb < t1; tb += n_split * kTG) {
      double tp[kTG][DIM], v[kTG][KN];
      load_targets(tb, tp, v);
      const int nt = min(kTG, t1 - tb);
      for (int q = lane; q < nc; q += 32) {
        double sp[DIM], w[KM];
#pragma unroll
        for (int a = 0; a < DIM; ++a) sp[a] = dat[a * kSrcCap + q];
#pragma unroll
        for (int m = 0; m < KM; ++m) w[m] = dat[(DIM + m) * kSrcCap + q];
        eval_source(sp, w, tp, v, nt);
      }
      store_targets(tb, nt, v);
    }
  }
  }
}

// ------------------------------------------------------------------------------------
// M2L operator tabulation.  For level l, cell width w, node spacing h = w / (order - 1) and
// cell offset o (source minus target, in cells) the operator is Toeplitz in the node index
// difference delta:  k(h * delta - w * o).  It is embedded in a circulant of length
// nf = 2 * order - 1 per axis and diagonalised:  Khat[o][b][a][f] = DFT(T)[f] / nf^dim.
// One CTA per offset; scratch in global memory (one-off precompute per configuration).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void dft_r2c_last(const double* in, double2* out, int outer, int nf, int p,
                                             const double2* __restrict__ tw) {
  const int total = outer * p;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    int kq = e % p, o = e / p;
    double re = 0.0, im = 0.0;
    int idx = 0;
    for (int n = 0; n < nf; ++n) {
      double v = in[static_cast<size_t>(o) * nf + n];
      double2 w = tw[idx];
      re = fma(v, w.x, re);
      im = fma(v, w.y, im);
      idx += kq;
      if (idx >= nf) idx -= nf;
    }
    out[e] = make_double2(re, im);
  }
}

__device__ __forceinline__ void dft_c2c_axis(const double2* in, double2* out, int outer, int nf, int inner,
                                             const double2* __restrict__ tw) {
  const int total = outer * nf * inner;
  for (int e = threadIdx.x; e < total; e += blockDim.x) {
    int i = e % inner, kq = (e / inner) % nf, o = e / (inner * nf);
    double re = 0.0, im = 0.0;
    int idx = 0;
    for (int n = 0; n < nf; ++n) {
      double2 v = in[(static_cast<size_t>(o) * nf + n) * inner + i];
      double2 w = tw[idx];
      re = fma(v.x, w.x, re);
      re = fma(-v.y, w.y, re);
      im = fma(v.x, w.y, im);
      im = fma(v.y, w.x, im);
      idx += kq;
      if (idx >= nf) idx -= nf;
    }
    out[e] = make_double2(re, im);
  }
}

// mode 0: the 7^dim child offsets, far ones only (the M2L lists);
// mode 1: only the 3^dim - 1 near child offsets, NEGATED (correction pass of the parent-block M2L, fmm_blk.cu);
// mode 2: the 3^dim - 1 parent offsets of the parent-block M2L (`it` then describes the block grid: order = 2p - 1
//         nodes, nf = 4p - 3, cell_w = the parent's width).
template <int FAM, int KIND, int DIM>
__global__ void __launch_bounds__(128) k_tabulate_m2l(RbfConst k, double cell_w, InterpDev it,
                                                      double2* __restrict__ Khat, double* scratch_t,
                                                      double2* scratch_c, int mode) {
  constexpr int KM = KindTraits<KIND, DIM>::km;
  constexpr int KN = KindTraits<KIND, DIM>::kn;
  const int p = it.order, nf = it.nf;
  const int base = mode == 2 ? 3 : 7, half = mode == 2 ? 1 : 3;
  int o3[DIM], r = blockIdx.x;
  bool near = true, zero = true;
#pragma unroll
  for (int a = DIM - 1; a >= 0; --a) {
    o3[a] = (r % base) - half;
    r /= base;
    near = near && o3[a] >= -1 && o3[a] <= 1;
    zero = zero && o3[a] == 0;
  }
  if (zero || (mode == 0 && near) || (mode == 1 && !near)) return;
  int NT = 1, F = p;
  for (int a = 0; a < DIM; ++a) NT *= nf;
  for (int a = 0; a + 1 < DIM; ++a) F *= nf;
  double* T = scratch_t + static_cast<size_t>(blockIdx.x) * NT;
  double2* bufA = scratch_c + static_cast<size_t>(blockIdx.x) * 2 * F;
  double2* bufB = bufA + F;
  const double h = cell_w / (p - 1);
  double scale = mode == 1 ? -1.0 : 1.0;
  for (int a = 0; a < DIM; ++a) scale /= nf;
  for (int comp = 0; comp < KN * KM; ++comp) {
    for (int e = threadIdx.x; e < NT; e += blockDim.x) {
      double d[DIM];
      int rr = e;
#pragma unroll
      for (int a = DIM - 1; a >= 0; --a) {
        int ia = rr % nf;
        rr /= nf;
        int delta = ia < p ? ia : ia - nf;
        // modes 1 and 2 take the distance from the INTEGER node offset, so that a node pair seen through the block grid
        // and through the child grid gets the same bits and coincident nodes of touching cells are at distance exactly 0
        // (kernels that are discontinuous there, e.g. the gradient of |r|, must cancel exactly between the two passes)
        d[a] = mode == 0 ? h * delta - cell_w * o3[a] : h * static_cast<double>(delta - (p - 1) * o3[a]);
      }
      double blk[KN * KM];
      kernel_block<FAM, KIND, DIM>(k, d, blk);
      T[e] = blk[comp] * scale;
    }
    __syncthreads();
    double2* dst = Khat + (static_cast<size_t>(blockIdx.x) * KN * KM + comp) * F;
    if constexpr (DIM == 1) {
      dft_r2c_last(T, dst, 1, nf, p, it.tw);
    } else {
      dft_r2c_last(T, bufA, NT / nf, nf, p, it.tw);
      __syncthreads();
      double2* in = bufA;
      double2* ob = bufB;
      int outer = NT / nf / nf;
      int inner = p;
      for (int a = DIM - 2; a >= 0; --a) {
        dft_c2c_axis(in, a == 0 ? dst : ob, outer, nf, inner, it.tw);
        __syncthreads();
        double2* t = in; in = ob; ob = t;
        inner *= nf;
        outer /= nf;
      }
    }
    __syncthreads();
  }
}

}  // namespace

void launch_p2p(int kind, int dim, const RbfConst& k, const TreeView& src, const double* swt, const TreeView& trg,
                double* vt, const int* leaves, int n_leaves, int64_t leaf_lo, int64_t leaf_hi, cudaStream_t s,
                LaunchCounter& c) {
  if (n_leaves <= 0 || leaf_hi <= leaf_lo) return;
  DevBuf<int> queue;  // stream-ordered pool allocation: no synchronisation
  queue.alloc(1, s);
  queue.zero(s);
  // heavy leaves (more than ~64 targets on average) are split across several queue items
  const int64_t avg_targets = trg.n / std::max(1, trg.n_cells[trg.height - 1]);
  const int n_split = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(16, avg_targets / 32)));
  const int grid = static_cast<int>(std::min<int64_t>(ceil_div(static_cast<int64_t>(n_leaves) * n_split, kP2PWarps),
                                                      num_sm() * 8));
  dispatch_fkd(k.family, kind, dim, [&](auto fam, auto knd, auto dm) {
    constexpr int km = KindTraits<knd.value, dm.value>::km;
    const size_t smem = sizeof(double) * kP2PWarps * (dm.value + km) * kSrcCap;
    if (smem > 40 * 1024)  // static shared memory counts towards the 48 KiB default limit
      PLT_CUDA(cudaFuncSetAttribute((const void*)k_p2p<fam.value, knd.value, dm.value>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    PLT_LAUNCH(c, (k_p2p<fam.value, knd.value, dm.value>), grid, kP2PWarps * 32, smem, s, k, src, swt, trg, vt, leaves,
               n_leaves, static_cast<int>(leaf_lo), static_cast<int>(leaf_hi), queue.get(), n_split);
  });
}

void launch_tabulate_m2l(int kind, int dim, const RbfConst& k, const Box& box, int level, const InterpDev& it,
                         double2* Khat_level, cudaStream_t s, LaunchCounter& c) {
  const int nf = it.nf;
  const int noff = ipow(7, dim);
  const size_t NT = ipow(nf, dim);
  const size_t F = freqs_per_cell(it.order, dim);
  DevBuf<double> st;
  DevBuf<double2> sc;
  st.alloc(noff * NT, s);
  sc.alloc(noff * 2 * F, s);
  const double cell_w = box.width / static_cast<double>(1 << level);
  dispatch_fkd(k.family, kind, dim, [&](auto fam, auto knd, auto dm) {
    PLT_LAUNCH(c, (k_tabulate_m2l<fam.value, knd.value, dm.value>), noff, 128, 0, s, k, cell_w, it, Khat_level,
               st.get(), sc.get(), 0);
  });
}

void launch_tabulate_m2l_blk(int kind, int dim, const RbfConst& k, const Box& box, int level, int order,
                             const double2* tw_blk, const double2* tw_child, double2* Kblk_level, double2* Khat_level,
                             cudaStream_t s, LaunchCounter& c) {
  const double cell_w = box.width / static_cast<double>(1 << level);
  InterpDev child{};  // only order, nf and tw are read by the tabulation
  child.order = order;
  child.nf = 2 * order - 1;
  child.tw = tw_child;
  InterpDev blk{};
  blk.order = blk_nodes(order);
  blk.nf = blk_nf(order);
  blk.tw = tw_blk;
  const int noff = ipow(7, dim), noff_b = ipow(3, dim);
  const size_t NT = ipow(child.nf, dim), F = freqs_per_cell(order, dim);
  const size_t NTB = ipow(blk.nf, dim), FB = blk_freqs(order);
  DevBuf<double> st;
  DevBuf<double2> sc;
  st.alloc(std::max(noff * NT, noff_b * NTB), s);
  sc.alloc(std::max(noff * 2 * F, noff_b * 2 * FB), s);
  dispatch_fkd(k.family, kind, dim, [&](auto fam, auto knd, auto dm) {
    PLT_LAUNCH(c, (k_tabulate_m2l<fam.value, knd.value, dm.value>), noff, 128, 0, s, k, cell_w, child, Khat_level,
               st.get(), sc.get(), 1);
    PLT_LAUNCH(c, (k_tabulate_m2l<fam.value, knd.value, dm.value>), noff_b, 128, 0, s, k, 2.0 * cell_w, blk, Kblk_level,
               st.get(), sc.get(), 2);
  });
}

}  // namespace plt
