// FMM operator kernels (device) -- host-callable launchers.
//
// These replace the ScalFMM passes the reference drives through
//   scalfmm::algorithms::fmm[omp](tree, op, p2m | m2m | m2l | l2l | l2p | p2p)
//   (src/fmm/fmm_evaluator.hpp:85-99, src/fmm/fmm_symmetric_evaluator.hpp:81-82).
#pragma once

#include <cuda_runtime.h>

#include "common.cuh"
#include "interp.hpp"
#include "rbf.cuh"
#include "tree.cuh"

namespace plt {

// Device copy of InterpTables.
struct InterpDev {
  int order;
  int nf;
  const double* beta;    // [order]
  const double* child;   // [2][order][order]
  const double2* tw;     // [nf] forward twiddles (cos, -sin)
  // Host copies (owned by the Interpolator) for kernels that take the tables as by-value
  // kernel parameters (constant bank): twiddles [nf][2], child [2][order][order], beta [order].
  const double* host_tw = nullptr;
  const double* host_child = nullptr;
  const double* host_beta = nullptr;
  bool polynomial = false;  // d == kClassic: the interpolant is the polynomial one (not Floater-Hormann)
};

inline int ipow(int b, int e) {
  int r = 1;
  for (int i = 0; i < e; ++i) r *= b;
  return r;
}
// Nodes per cell and half-spectrum size per cell.
inline int nodes_per_cell(int order, int dim) { return ipow(order, dim); }
inline int freqs_per_cell(int order, int dim) { return ipow(2 * order - 1, dim - 1) * order; }

// ---- point pre/post processing ----
// pos_out[a][i] = sum_b A[a][b] * points[i][b]   (caller order, SoA)
void launch_transform_points(int dim, const double* aniso, const double* points_rowmajor, int64_t n,
                             double* pos_soa, cudaStream_t s, LaunchCounter& c);
// wt[m][i] = folded weights of sorted point i  (perm == nullptr: caller order)
void launch_prepare_weights(int kind, int dim, const double* aniso, const double* weights, const int* perm,
                            int64_t n, double* wt_soa, cudaStream_t s, LaunchCounter& c);
// out[kn*perm[i] + b] = (A^T applied if kind in {FT, H}) vt[.][i]  (+ self term k(0) w_i)
void launch_finish_outputs(int kind, int dim, const double* aniso, const double* vt_soa, const int* perm,
                           int64_t n, int64_t lo, int64_t hi, double* out, cudaStream_t s, LaunchCounter& c);

// ---- upward ----
void launch_p2m(int dim, int km, const TreeView& tr, const Box& box, const InterpDev& it, const double* wt,
                double* M, cudaStream_t s, LaunchCounter& c);
void launch_m2m(int dim, int km, const TreeView& tr, int parent_level, const InterpDev& it, double* M,
                cudaStream_t s, LaunchCounter& c);
// M (real nodes) -> Mhat (half spectrum) for all cells of levels [2, height).
void launch_m2hat(int dim, int km, const TreeView& tr, const InterpDev& it, const double* M, double2* Mhat,
                  cudaStream_t s, LaunchCounter& c);

// ---- M2L ----
// Tabulate the Fourier-space M2L operators of one level: Khat[oi][b][a][f], oi over 7^dim offsets.
void launch_tabulate_m2l(int kind, int dim, const RbfConst& k, const Box& box, int level, const InterpDev& it,
                         double2* Khat_level, cudaStream_t s, LaunchCounter& c);

struct M2LArgs {
  TreeView trg;
  int level;             // target/source cell level (>= 2)
  int order, dim, km, kn;
  const double2* Mhat;   // all source cells, indexed by global compact id - cell_off[2]
  const double2* Khat;   // this level's operators
  // Slices of the plan (plan.cuh) for the chunk of active parents being processed:
  const int* active;     // [n_active] compact ids (level-1 of the target tree)
  const int* src_ids;    // [n_active][3^dim * 2^dim] Mhat cell index or -1
  const unsigned char* trg_mask;  // [n_active] existing target children
  int n_active;
  double2* Lhat;         // scratch [n_active][2^dim][kn][F]
  double* L;             // target locals of the upper levels, indexed by global compact id (or null)
  double* Lc;            // compact leaf-level output [n_active][2^dim][kn][P] (used when L == null)
};
// Fourier-space accumulation over the M2L lists of the children of the active parents.
void launch_m2l_hadamard(const M2LArgs& a, cudaStream_t s, LaunchCounter& c);
// counters[0] += M2L pairs, [1] += target cells with a non-empty M2L list, [2] += P2P pairs.
void launch_count_work(int dim, const TreeView& src, const TreeView& trg, unsigned long long* counters,
                       cudaStream_t s, LaunchCounter& c);
void launch_m2l_idft(const M2LArgs& a, const InterpDev& it, cudaStream_t s, LaunchCounter& c);

// ---- downward ----
// children [cell_lo, cell_hi) of level child_level from their parents [par_lo, par_hi) of level child_level - 1
void launch_l2l(int dim, int kn, const TreeView& tr, int child_level, const InterpDev& it, double* L,
                int cell_lo, int cell_hi, int par_lo, int par_hi, cudaStream_t s, LaunchCounter& c);
void launch_l2p(int dim, int kn, const TreeView& tr, const Box& box, const InterpDev& it, const double* L,
                double* vt, int64_t lo, int64_t hi, cudaStream_t s, LaunchCounter& c);
// Fused last level of the downward pass: for every parent of level leaf-1, L2L to its children
// in shared memory (+ the children's own M2L result Lc, slot given by the plan's leaf_meta), then L2P.  The
// leaf-level local expansions never touch HBM.  Returns false when the order is too large for
// the shared-memory staging (caller falls back to launch_l2l + launch_l2p).
bool launch_l2l_l2p_leaf(int dim, int kn, const TreeView& tr, const Box& box, const InterpDev& it, const double* L,
                         const double* Lc, const int* leaf_meta, double* vt, int64_t leaf_lo, int64_t leaf_hi,
                         int par_lo, int par_hi, cudaStream_t s, LaunchCounter& c);
size_t leaf_fused_smem_bytes(int dim, int order);
bool leaf_fused_supported(int dim, int order);
// Near field over the 3^dim adjacent source leaves of the listed target leaves (ascending ids,
// restricted to [lo, hi)); vt += ...
void launch_p2p(int kind, int dim, const RbfConst& k, const TreeView& src, const double* swt, const TreeView& trg,
                double* vt, const int* leaves, int n_leaves, int64_t lo, int64_t hi, cudaStream_t s,
                LaunchCounter& c);

}  // namespace plt
