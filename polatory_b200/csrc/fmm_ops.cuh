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
// pos_out[a][i] = sum_b A[a][b] * points[i][b]   (caller order, SoA; rows `ld` apart, ld = 0: n)
void launch_transform_points(int dim, const double* aniso, const double* points_rowmajor, int64_t n,
                             double* pos_soa, cudaStream_t s, LaunchCounter& c, int64_t ld = 0);
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
  // Parent-block M2L (3-D, fmm_blk.cu); unused (null) on the list path.
  const double2* Mblk = nullptr;   // block spectra of all source cells of levels 1 .. height-2 (id - cell_off[1])
  const double2* Kblk = nullptr;   // this level's block operators [27][kn][km][FB]
  const int* grp_first = nullptr;  // plan: first slot of each sibling group of active parents (all levels, ascending)
  const int* grp_slot = nullptr;   // plan: [group][8] slot of the parent at sibling position tp, or -1
  const int* grp_src = nullptr;    // plan: [group][64] Mblk row of the source parent at position sp of the 4^3 block, or -1
  int grp_lo = 0, grp_hi = 0;      // this level's group range
  int slot0 = 0;                   // plan slot of active[0] (groups are cut to the chunk [slot0, slot0 + n_active))
  double2* Lhat_blk = nullptr;     // scratch [n_active][kn][FB]
  // Several consecutive levels in ONE Hadamard launch (the levels above the leaf level are a few launches of a few
  // hundred parents each, latency-bound one by one): slots [lvl_slot_end[i-1], lvl_slot_end[i]) of the chunk belong to
  // level `level + i`, whose operators are Khat + i * khat_level_stride.  n_lvls = 1: the chunk is one level.
  int n_lvls = 1;
  int lvl_slot_end[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  size_t khat_level_stride = 0;
};
// Fourier-space accumulation over the M2L lists of the children of the active parents.
void launch_m2l_hadamard(const M2LArgs& a, cudaStream_t s, LaunchCounter& c);
// May a chunk span several levels (M2LArgs::n_lvls > 1)?  (the tiled list kernel only)
bool m2l_hadamard_multi_level_supported();
// Scalar 3-D variant with the operators of most pairs in Tensor Memory (fmm_had_tmem.cu); false = not applicable.
bool launch_m2l_hadamard_tmem(const M2LArgs& a, int F, cudaStream_t s, LaunchCounter& c);
bool hadamard_tmem_enabled();
int hadamard_tmem_set(int on);  // returns the previous setting
// counters[0] += M2L pairs, [1] += target cells with a non-empty M2L list, [2] += P2P pairs.
void launch_count_work(int dim, const TreeView& src, const TreeView& trg, unsigned long long* counters,
                       cudaStream_t s, LaunchCounter& c);
void launch_m2l_idft(const M2LArgs& a, const InterpDev& it, cudaStream_t s, LaunchCounter& c);

// ---- M2L by parent blocks (3-D; fmm_blk.cu) ----
// The equispaced nodes of the 8 children of a cell form ONE equispaced grid of (2 order - 1)^3 distinct nodes (the
// children share their boundary planes).  The interaction of all children of a source parent with all children of a
// neighbouring target parent is therefore a single Toeplitz operator on that grid, diagonalised by a DFT of length
// 4 order - 3 per axis: one complex multiply-add per frequency and parent pair instead of one per CHILD pair and child
// frequency (9.6 x fewer at order 6).  That sum also contains the ADJACENT child pairs of different parents, which do
// not belong to the M2L list; they are subtracted by the child-level Hadamard kernel run over the near offsets with
// negated operators (launch_m2l_hadamard_near).  Exact up to rounding; needs k finite at distance 0 (touching cells
// have coincident nodes), which the caller checks on the tabulated operators.
inline int blk_nodes(int order) { return 2 * order - 1; }
inline int blk_nf(int order) { return 4 * order - 3; }
inline size_t blk_freqs(int order) { return static_cast<size_t>(blk_nf(order)) * blk_nf(order) * blk_nodes(order); }
bool blk_supported(int dim, int order);
// Process-wide density threshold (plt_set_block_m2l_min_fill) and the per-level decision: the children level l goes
// through the block path when the source tree has at least min_fill * 8^l cells there.
double blk_min_fill();
double blk_set_min_fill(double v);
// Small levels stay on the list path unless the block path is forced (min_fill <= 0): their launches are
// latency-bound either way and the extra passes do not pay (measured on config #3: +0.4 ms for levels 2 - 4).
constexpr int kBlkMinCells = 8192;
inline bool blk_level_dense(int level, int n_src_cells) {
  const double f = blk_min_fill();
  return static_cast<double>(n_src_cells) >= f * static_cast<double>(1ll << (3 * level)) &&
         (f <= 0.0 || n_src_cells >= kBlkMinCells);
}
// Block operators of one level: Kblk[D][b][a][f], D over the 3^3 parent offsets (centre unused), and the negated
// child-level operators of the 3^3 near offsets written into Khat_level (which the far tabulation leaves at zero).
void launch_tabulate_m2l_blk(int kind, int dim, const RbfConst& k, const Box& box, int level, int order,
                             const double2* tw_blk, const double2* tw_child, double2* Kblk_level, double2* Khat_level,
                             cudaStream_t s, LaunchCounter& c);
// flag[0] |= 1 if any of the n doubles is not finite.
void launch_check_finite(const double* x, size_t n, int* flag, cudaStream_t s, LaunchCounter& c);
// Block spectra of all source cells of levels 1 .. height-2 from the multipoles of their children.
// (parents whose level is in [par_lo, par_hi] only; Mblk is indexed by global cell id - cell_off[1])
void launch_mblk(int km, const TreeView& src, int order, int par_lo, int par_hi, const double* M, double2* Mblk,
                 cudaStream_t s, LaunchCounter& c);
void launch_m2l_blk_hadamard(const M2LArgs& a, cudaStream_t s, LaunchCounter& c);
void launch_m2l_hadamard_near(const M2LArgs& a, cudaStream_t s, LaunchCounter& c);
// L (or Lc) of the children += pruned inverse DFT of Lhat_blk; runs after launch_m2l_idft (which stores).
void launch_m2l_blk_idft(const M2LArgs& a, cudaStream_t s, LaunchCounter& c);

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
