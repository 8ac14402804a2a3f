/*
 * polatory_b200 -- C ABI of the B200-native RBF fast-multipole evaluator.
 *
 * This is the drop-in boundary for Polatory's src/fmm evaluator interface
 * (SURVEY.md section 8b).  One handle replaces one instance of
 *
 *   polatory::fmm::FmmGenericEvaluatorBase<Dim>          include/polatory/fmm/fmm_evaluator.hpp:17-41
 *   polatory::fmm::FmmGenericSymmetricEvaluatorBase<Dim>  include/polatory/fmm/fmm_symmetric_evaluator.hpp:16-37
 *
 * as returned by the six factories
 *
 *   make_fmm_evaluator / make_fmm_gradient_evaluator / make_fmm_gradient_transpose_evaluator /
 *   make_fmm_hessian_evaluator                            include/polatory/fmm/fmm_evaluator.hpp:92-106
 *   make_fmm_symmetric_evaluator / make_fmm_hessian_symmetric_evaluator
 *                                                         include/polatory/fmm/fmm_symmetric_evaluator.hpp:80-86
 *
 * Conventions (same as the reference, SURVEY.md 8b "Data conventions"):
 *   - points:  contiguous row-major N x dim doubles in ORIGINAL coordinates; the
 *              evaluator applies the anisotropy (src/fmm/fmm_evaluator.hpp:120-157);
 *   - weights: km doubles per source point, point-major (km*idx + i);
 *   - result:  kn doubles per target point, point-major, in caller order;
 *   - km/kn:   K 1/1, F dim/1, FT 1/dim, H dim/dim
 *              (include/polatory/fmm/kernel.hpp:29-30, gradient_kernel.hpp:28-29,
 *               gradient_transpose_kernel.hpp:29-30, hessian_kernel.hpp:28-29).
 *   - every pointer argument may be a host pointer OR a device pointer of the
 *     handle's device (CUDA unified addressing decides); data are copied in, nothing is
 *     borrowed after the call returns.  Work is issued on the handle's stream
 *     (default: the legacy default stream); calls taking host output pointers
 *     synchronise that stream before returning.
 *   - no exceptions cross this boundary: every call returns a status code and
 *     plt_last_error() gives the message a C++ shim rethrows as std::runtime_error.
 *   - there is NO CPU fallback: without a CUDA device every call fails with
 *     PLT_ERR_CUDA.
 */
#ifndef POLATORY_B200_H_
#define POLATORY_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct plt_eval plt_eval;

/* Status codes. */
enum {
  PLT_OK = 0,
  PLT_ERR_INVALID = 1,   /* bad argument (unknown RBF, bad dim, size mismatch, ...)        */
  PLT_ERR_CUDA = 2,      /* CUDA runtime failure (incl. "no device")                        */
  PLT_ERR_ACCURACY = 3,  /* "failed to construct an evaluator that meets the desired
                            accuracy"  src/fmm/fmm_accuracy_estimator.hpp:120               */
  PLT_ERR_UNSUPPORTED = 4 /* e.g. Hessian of cov_spherical / cov_cubic
                            include/polatory/rbf/cov_spherical.hpp:53-55                    */
};

/* Kernel kinds = the four kernel functors of include/polatory/fmm/. */
enum { PLT_KIND_K = 0, PLT_KIND_F = 1, PLT_KIND_FT = 2, PLT_KIND_H = 3 };

/* RBF ids = the 16 runtime RBFs of include/polatory/rbf/make_rbf.hpp:30-45. */
enum {
  PLT_RBF_BH3 = 0, PLT_RBF_TH3 = 1, PLT_RBF_BH2 = 2, PLT_RBF_TH2 = 3,
  PLT_RBF_EXP = 4, PLT_RBF_GAU = 5,
  PLT_RBF_GC3 = 6, PLT_RBF_GC5 = 7, PLT_RBF_GC7 = 8, PLT_RBF_GC9 = 9,
  PLT_RBF_SP3 = 10, PLT_RBF_SP5 = 11, PLT_RBF_SP7 = 12, PLT_RBF_SP9 = 13,
  PLT_RBF_SPH = 14, PLT_RBF_CUB = 15
};

/* Spheroidal split (include/polatory/rbf/cov_spheroidal3.hpp:113-123): the full RBF is
 * evaluated as direct part (compact) + fast part (FMM), src/fmm/spheroidal_evaluator.hpp:24-29.
 * PLT_PART_FULL on a spheroidal id performs that split internally. */
enum { PLT_PART_FULL = 0, PLT_PART_DIRECT = 1, PLT_PART_FAST = 2 };

/* Interpolator configuration, src/fmm/interpolator_configuration.hpp:9-19. */
typedef struct plt_config {
  int tree_height; /* 0 => brute-force branch (src/fmm/fmm_evaluator.hpp:226-234) */
  int order;
  int d;           /* -1 = classic polynomial; 0..order-1 = Floater-Hormann degree */
} plt_config;

/* Replaces the six factories: ctor(rbf, bbox) of FmmGenericEvaluator<Kernel> /
 * FmmGenericSymmetricEvaluator<Kernel> (src/fmm/fmm_evaluator.hpp:71-76,
 * src/fmm/fmm_symmetric_evaluator.hpp:59-64).  params: the RBF's parameters()
 * (n_params 0..2; polyharmonics default to {1, 0}, polyharmonic_odd.hpp:87-99);
 * aniso: dim*dim row-major anisotropy matrix (NULL = identity);
 * bbox_min/max: dim doubles each, must contain all sources and targets. */
int plt_eval_create(int kind, int symmetric, int dim, int rbf_id, int rbf_part,
                    const double* params, int n_params, const double* aniso,
                    const double* bbox_min, const double* bbox_max, plt_eval** out);

void plt_eval_destroy(plt_eval* h);

/* FmmGenericEvaluatorBase::set_source_points / set_target_points
 * (include/polatory/fmm/fmm_evaluator.hpp:36-38).  Invalid on a symmetric handle. */
int plt_eval_set_source_points(plt_eval* h, const double* points, int64_t n);
int plt_eval_set_target_points(plt_eval* h, const double* points, int64_t n);

/* FmmGenericSymmetricEvaluatorBase::set_points (fmm_symmetric_evaluator.hpp:33). */
int plt_eval_set_points(plt_eval* h, const double* points, int64_t n);

/* set_weights (fmm_evaluator.hpp:40): len must be km * n_sources. */
int plt_eval_set_weights(plt_eval* h, const double* weights, int64_t len);

/* set_accuracy (fmm_evaluator.hpp:34): +inf => order 6 polynomial, 0 => order 12 / d 8,
 * finite => search order 8,10,...,20 (src/fmm/fmm_accuracy_estimator.hpp:74-121). */
int plt_eval_set_accuracy(plt_eval* h, double accuracy);

/* evaluate (fmm_evaluator.hpp:32): writes kn * n_targets doubles. */
int plt_eval_evaluate(plt_eval* h, double* out, int64_t len);

/* set_target_points(points) + evaluate(out) in one call: the sequence of interpolation::Evaluator::evaluate(points)
 * (include/polatory/interpolation/evaluator.hpp:83-87).  With HOST buffers (pinned ones for full overlap) on the FMM
 * branch the targets are streamed in caller-order slabs through a copy-in / evaluate / copy-out pipeline on three
 * streams; every slab is evaluated against the octree of the whole problem, so the values are bit-identical to the two
 * separate calls.  Device pointers, small problems, symmetric / partitioned / compact-support evaluators take the two
 * calls as they are.  On return the evaluator is in the state the two calls leave it in. */
int plt_eval_evaluate_points(plt_eval* h, const double* points, int64_t n, double* out, int64_t len);

/* Additions with no reference counterpart. */

/* Bypass the accuracy search with a fixed (order, d); order 0 restores the search.
 * tree_height is always derived from the point counts (src/fmm/utility.hpp:12-16)
 * unless tree_height_override > 0. */
int plt_eval_force_config(plt_eval* h, int order, int d, int tree_height_override);

/* Always evaluate by exact direct summation (the brute-force branch of src/fmm/fmm_evaluator.hpp:226-234
 * regardless of the problem size): what interpolation::DirectEvaluator does for the <= 1024 sampled targets
 * of ResidualEvaluator (include/polatory/interpolation/residual_evaluator.hpp:55-88). */
int plt_eval_force_direct(plt_eval* h, int on);

/* The configuration used by the last evaluate() (tree_height 0 = brute force). */
int plt_eval_get_config(plt_eval* h, plt_config* out);

/* Issue all work of this handle on the given cudaStream_t (NULL = legacy default). */
int plt_eval_set_stream(plt_eval* h, void* cuda_stream);

/* Restrict evaluate() to the targets [begin, end) of the Morton-sorted target order
 * (multi-GPU sharding by Morton range); outputs of other targets are
 * written as 0 so that a sum over ranks reassembles the full result.
 * begin = 0, end = -1 restores the full range. */
int plt_eval_set_target_shard(plt_eval* h, int rank, int world_size);

/* Multi-GPU evaluation across one node, one process per GPU.  Every rank holds all source points
 * and weights (24 + 8 bytes per source); the level-`level_cut` cells of the octree are partitioned by Morton
 * key: rank r OWNS the keys [key_begin[r], key_begin[r + 1]) (key_begin[0] = 0; key_begin[world_size] is
 * taken as 2^(dim * level_cut)).
 *   - A generic evaluator is then given only this rank's targets (ideally those of its own key range);
 *     call plt_eval_force_config(h, 0, -1, height) with the height of the GLOBAL problem
 *     (src/fmm/utility.hpp:12-16 on the global point counts) so that every rank builds the same octree.
 *   - A symmetric evaluator (the matvec) is given all points and evaluates the leaves of its own key range;
 *     outputs of the other points are written as 0 (plt_eval_get_target_shard_range gives the point range).
 * The upward pass is partitioned: a rank computes P2M / M2M / the multipole spectra only below the level-cut
 * cells it owns or that lie within the interaction range of its targets; the level-cut expansions are then
 * exchanged with ONE all-gather (`allgatherv`, e.g. NCCL over NVLink) and the few upper levels are finished
 * on every rank.  Results are bit-identical to the single-GPU evaluation of the same targets.
 * world_size <= 1 removes the partition.
 *
 * allgatherv(ctx, buf, offsets, world_size, stream): in-place all-gather on DEVICE memory; rank r contributes the
 * doubles [offsets[r], offsets[r + 1]) of `buf`; on return every segment is filled.  Work must be ordered after
 * / before the other work of `stream` (a cudaStream_t).  Returns 0 on success. */
typedef int (*plt_allgatherv_fn)(void* ctx, double* buf, const int64_t* offsets, int world_size, void* stream);
int plt_eval_set_partition(plt_eval* h, int rank, int world_size, int level_cut, const uint32_t* key_begin,
                           plt_allgatherv_fn allgatherv, void* ctx);
/* src/fmm/utility.hpp:12-16: max(2, round(ln n / ln 2^dim)); n = max(n_src, n_trg) (symmetric: n). */
int plt_tree_height(int dim, int64_t n_points);
/* Morton keys at `level` of host points (original coordinates; the evaluator's anisotropy and root box are
 * applied): what a caller partitions its targets with.  keys: n uint32. */
int plt_eval_point_keys(plt_eval* h, const double* points, int64_t n, int level, uint32_t* keys);
/* Number of all-gathers issued by this handle so far. */
int64_t plt_eval_allgather_count(plt_eval* h);

/* Multi-GPU matvec support: the permutation of the target tree, perm[i] =
 * caller index of the i-th point in Morton order (the order target shards are cut in), and
 * the sorted-order point range [begin, end) of the current target shard (cuts are moved to
 * leaf boundaries).  A caller that feeds its points already in this order gets an identity
 * permutation (the sort is stable), so each rank's shard is a contiguous slice of the
 * Krylov vectors.  Both build the tree(s) if needed; on the brute-force branch (no tree) the
 * permutation is the identity and rank 0 owns everything. */
int plt_eval_get_permutation(plt_eval* h, int32_t* perm, int64_t n);
int plt_eval_get_target_shard_range(plt_eval* h, int64_t* begin, int64_t* end);

/* Batched Gram matrices of the value kernel for the RAS preconditioner's local problems: replaces
 * preconditioner::mat_a (include/polatory/preconditioner/mat_a.hpp:10-61, value block) for all domains
 * of a level at once.  points: DEVICE [n_batch][m][dim] row-major, original coordinates (the handle's
 * anisotropy is applied); counts: DEVICE int32 [n_batch], rows / columns >= counts[b] are filled with the
 * identity; out: DEVICE [n_batch][m][m].  out[b][i][j] = phi(x_bi - x_bj) + nugget * (i == j).
 * The handle must be a K-kind evaluator of the RBF (it carries the constants); n_batch <= 65535. */
int plt_eval_gram_batched(plt_eval* h, const double* points, const int32_t* counts, int64_t n_batch, int64_t m,
                          double nugget, double* out);

/* The same for Hermite data (the F / F^T / H blocks of mat_a): one row per value point and `dim` rows per
 * gradient point.  types: DEVICE int8 [n_batch][m]: 0 = value row, 1 + c = gradient component c, negative =
 * padding (identity); points: the coordinates of each ROW's point. */
int plt_eval_gram_mixed(plt_eval* h, const double* points, const int8_t* types, int64_t n_batch, int64_t m,
                        double nugget, double* out);

/* Per-phase device time of the last evaluate() in milliseconds (CUDA events on the
 * handle's stream).  names/ms: arrays of capacity cap; returns the number of phases. */
int plt_eval_phase_times(plt_eval* h, const char** names, double* ms, int cap);

/* Algorithmic work of the current (source tree, target tree) pair, counted on the device:
 * M2L cell pairs, target cells with a non-empty M2L list (all levels), P2P point pairs.
 * Diagnostic (feeds the roofline figures); valid after an evaluate() on the FMM branch. */
int plt_eval_work_stats(plt_eval* h, int64_t* m2l_pairs, int64_t* m2l_target_cells, int64_t* p2p_pairs);

/* Number of kernels launched by this handle since creation. */
int64_t plt_eval_launch_count(plt_eval* h);

/* Message of the last failure on this handle (h == NULL: last plt_eval_create failure
 * on the calling thread).  Never NULL. */
const char* plt_last_error(plt_eval* h);

/* M2L by parent blocks (3-D, orders 6 and 8; polatory_b200/csrc/fmm_blk.cu): a level of the far field is computed
 * through the block spectra of sibling cells when the source tree fills at least `min_fill` of the cells of that
 * level and that level has at least 8192 source cells (dense tables: volume clouds), else through the per-pair
 * lists.  Process-wide; default 0.25; 0 = wherever supported, > 1 = never.  Both paths evaluate the same sums (rounding
 * differs); the choice depends on the source tree only, so every rank of a partition takes the same one.
 * Returns the previous value. */
double plt_set_block_m2l_min_fill(double min_fill);

/* Work-space blocks of destroyed evaluators are kept in a process-wide cache (up to a quarter of the device memory by
 * default) and reused by the next evaluators that ask for about the same size: a second fit or a second sampler in the
 * same process then performs no cudaMalloc / cudaFree of multi-GB blocks.  plt_release_cached_memory returns the cache to
 * the driver (bytes freed; also done automatically when an allocation fails), plt_cached_memory reports its size,
 * plt_set_cached_memory_limit(0) switches it off. */
int64_t plt_release_cached_memory(void);
int64_t plt_cached_memory(void);
void plt_set_cached_memory_limit(int64_t bytes);

/* Experimental: scalar 3-D Hadamard M2L with the operators of most pairs in Tensor Memory (tcgen05.ld; csrc/fmm_had_tmem.cu).
 * Off by default (parity-green, but slower than the shared-memory kernel on B200: profiles/r02_k_hadamard_tmem.md);
 * PLT_HAD_TMEM=1 in the environment or this switch turn it on.  Returns the previous setting. */
int plt_set_hadamard_tmem(int on);

/* FP64 FMA peak of the current device, measured with a DFMA-chain microbenchmark (TFLOP/s);
 * the denominator of the FP64 roofline. */
int plt_measure_fp64_peak(double* tflops);

/* ---------------------------------------------------------------------------------------
 * Device-resident flexible GMRES: replaces krylov::Fgmres
 *   include/polatory/krylov/gmres_base.hpp:11-91, src/krylov/gmres_base.cpp:7-85,
 *   src/krylov/gmres.cpp:9-50, src/krylov/fgmres.cpp:8-28
 * as driven by interpolation::Solver::solve (include/polatory/interpolation/solver.hpp:99-139).
 * Vectors stay in HBM; the operator / right preconditioner are callbacks receiving DEVICE
 * pointers of length n_local that must issue their work on the solver's stream and return 0.
 * With sharded vectors (one rank per GPU) the `allreduce` callback sums `count` doubles of a
 * DEVICE buffer in place across ranks (NCCL), which reduces the Krylov dot products.
 * ------------------------------------------------------------------------------------- */
typedef struct plt_fgmres plt_fgmres;
typedef int (*plt_linop_fn)(void* ctx, const double* x_dev, double* y_dev);
typedef int (*plt_allreduce_fn)(void* ctx, double* buf_dev, int count);

/* Fgmres(op, rhs, max_iter): n_local = local vector length, max_iter = maximum iterations
 * (no restart, as the reference). */
int plt_fgmres_create(int64_t n_local, int max_iter, plt_fgmres** out);
void plt_fgmres_destroy(plt_fgmres* h);
int plt_fgmres_set_operator(plt_fgmres* h, plt_linop_fn op, void* ctx);
/* set_right_preconditioner (gmres_base.cpp:33-37); NULL = identity.  Before setup(). */
int plt_fgmres_set_right_preconditioner(plt_fgmres* h, plt_linop_fn pc, void* ctx);
int plt_fgmres_set_allreduce(plt_fgmres* h, plt_allreduce_fn fn, void* ctx);
int plt_fgmres_set_stream(plt_fgmres* h, void* cuda_stream);
/* set_initial_solution + setup (gmres_base.cpp:27-31,39-51): rhs, x0 host or device pointers
 * (x0 NULL = zero). */
int plt_fgmres_setup(plt_fgmres* h, const double* rhs, const double* x0);
/* iterate_process (gmres.cpp:9-50): one Arnoldi step + Givens update; one host sync. */
int plt_fgmres_iterate(plt_fgmres* h);
/* solution_vector (fgmres.cpp:8-26): x = x0 + Z y; x host or device pointer. */
int plt_fgmres_solution(plt_fgmres* h, double* x);
/* iteration_count / absolute_residual / relative_residual (gmres_base.cpp:7-15). */
int plt_fgmres_status(plt_fgmres* h, int* iteration_count, double* absolute_residual, double* relative_residual);
int64_t plt_fgmres_launch_count(plt_fgmres* h);
const char* plt_fgmres_last_error(plt_fgmres* h);

/* ---------------------------------------------------------------------------------------
 * Host-side index bookkeeping of the RAS preconditioner (no device work, multi-threaded):
 *   DomainDivider::choose_coarse_points   include/polatory/preconditioner/domain_divider.hpp:52-123
 *   DomainDivider::divide_domains + Domain::merge_poly_points
 *                                         include/polatory/preconditioner/domain_divider.hpp:171-286, domain.hpp:33-51
 * for value points.  a_points: HOST row-major [*][dim] anisotropy-transformed coordinates, addressed by the
 * global indices in idcs / poly.
 * ------------------------------------------------------------------------------------- */
/* out: n_poly + n_coarse global indices (the poly points first, then the cluster centres). */
int plt_ras_choose_coarse_points(const double* a_points, int dim, const int64_t* idcs, int64_t n_idcs,
                                 const int64_t* poly, int64_t n_poly, int64_t n_coarse, int64_t* out);
typedef struct plt_ras_domains plt_ras_domains;
/* Recursive bisection into overlapping domains of at most max_leaf points (+ the poly points, which are
 * put first in every domain); reference constants: max_leaf 1024, overlap_quota 0.5. */
int plt_ras_divide_domains(const double* a_points, int dim, const int64_t* idcs, int64_t n_idcs, const int64_t* poly,
                           int64_t n_poly, int64_t max_leaf, double overlap_quota, plt_ras_domains** out);
/* The same two for Hermite data: value points and gradient points (multiplicity dim) mixed, with the reference's
 * multiplicity-weighted cut ranks (domain_divider.hpp:66-90, 205-231).  n_coarse_rows counts rows (a gradient
 * centre counts dim).  out_points needs n_poly + n_coarse_rows entries, out_grads n_coarse_rows. */
int plt_ras_choose_coarse_points_mixed(const double* a_points, const double* a_grad_points, int dim,
                                       const int64_t* point_idcs, int64_t n_points, const int64_t* grad_idcs,
                                       int64_t n_grads, const int64_t* poly, int64_t n_poly, int64_t n_coarse_rows,
                                       int64_t* out_points, int64_t* n_out_points, int64_t* out_grads,
                                       int64_t* n_out_grads);
int plt_ras_divide_domains_mixed(const double* a_points, const double* a_grad_points, int dim, const int64_t* point_idcs,
                                 int64_t n_points, const int64_t* grad_idcs, int64_t n_grads, const int64_t* poly,
                                 int64_t n_poly, int64_t max_leaf, double overlap_quota, plt_ras_domains** out);
int64_t plt_ras_domains_total_grads(plt_ras_domains* h);
int plt_ras_domains_get_grads(plt_ras_domains* h, int64_t* offsets, int64_t* indices, uint8_t* inner);
int64_t plt_ras_domains_count(plt_ras_domains* h);
int64_t plt_ras_domains_total(plt_ras_domains* h);
/* offsets: count + 1 entries; indices / inner: total entries (inner = 1 where the domain owns the point). */
int plt_ras_domains_get(plt_ras_domains* h, int64_t* offsets, int64_t* indices, uint8_t* inner);
void plt_ras_domains_destroy(plt_ras_domains* h);

/* Dense local problems of the RAS preconditioner (preconditioner/fine_grid.hpp:59-142, coarse_grid.hpp:40-131),
 * batched over the domains of a level, DEVICE pointers, row-major, hand-written kernels (no cuSOLVER / cuBLAS).
 *   plt_ras_reduce_q:        red[b] = Q^T A Q = A_rr + Q_top^T (A_tt Q_top + A_tr) + A_rt Q_top; a: [B][m][m] (the l
 *                            polynomial rows first), q_top: [B][l][m - l], red: [B][m - l][m - l]; l = 0: not needed.
 *   plt_chol_batched:        in-place Cholesky L L^T of [B][n][n] (lower triangle overwritten by L); info[b] = 0 or
 *                            1 + the first non-positive pivot (info may be NULL).
 *   plt_chol_solve_batched:  per domain qtd = vals[l:] + Q_top^T vals[:l]; L L^T gamma = qtd;
 *                            lam = [Q_top gamma; gamma]  (fine_grid.hpp:112-133); vals, lam: [B][l + n]. */
int plt_ras_reduce_q(const double* a, const double* q_top, int64_t n_batch, int m, int l, double* red, void* stream);
int plt_chol_batched(double* a, int64_t n_batch, int n, int* info, void* stream);
int plt_chol_solve_batched(const double* factor, int64_t n_batch, int n, const double* q_top, int l, const double* vals,
                           double* lam, void* stream);
/* ONE factor (as written by plt_chol_batched), n_rhs right-hand sides: out[k] = (L L^T)^-1 rhs[k]; rhs, out: [n_rhs][n].
 * With the identity as right-hand sides this forms the explicit inverse of the coarse grid's matrix once per fit. */
int plt_chol_solve_shared(const double* factor, int n, int64_t n_rhs, const double* rhs, double* out, void* stream);
/* y = A x, A row-major [rows][cols], device pointers (applies the coarse grid's inverse, coarse_grid.hpp:84-128). */
int plt_gemv(const double* a, int rows, int cols, const double* x, double* y, void* stream);

/* RasPreconditioner::operator() (include/polatory/preconditioner/ras_preconditioner.hpp:183-246) as ONE call on a
 * stream: the multiplicative level sweep -- coarse-grid solves (coarse_grid.hpp:84-128), batched fine-level solves
 * (fine_grid.hpp:103-147), update_residuals through resident evaluators (:287-321), orthogonalize (:267-285).
 * The handle references (does not own) DEVICE tables set up through the entry points above:
 *   set_level_rows: rows of the global system [mu values | dim * sigma gradient components] of a level's value points and,
 *                   point-major, of the components of its gradient points;
 *   set_fine:       a level >= 1: idx [n_domains][m] rows of every domain (the l polynomial rows first, padded), cnt
 *                   valid rows per domain, the factors written by plt_chol_batched [n_domains][m-l][m-l], q_top
 *                   [n_domains][l][m-l] (NULL for l = 0), and the (row, position in the [n_domains][m] solution) pairs
 *                   of the points each domain owns;
 *   set_coarse:     level 0: rows idx[m], the explicit inverse of Q^T A Q [(m-l)^2], q_top [l][m-l], the first l rows of
 *                   mat_a [l][m], the inverse of the polynomial matrix at the polynomial points [l][l];
 *   add_transfer:   an evaluator (kind 0 = K, 1 = F, 2 = F^T, 3 = H) whose sources are src_level's points and whose
 *                   targets are trg_level's points, both already set (several per pair: one model with several RBFs);
 *   set_poly:       monomials, orthonormalised monomials and A p at every row, [m_rows][l] row-major;
 *   apply:          out = M^-1 v, v and out: m_rows + l doubles on the device. */
typedef struct plt_ras_sweep plt_ras_sweep;
int plt_ras_sweep_create(int64_t m_rows, int l, int n_levels, plt_ras_sweep** out);
void plt_ras_sweep_destroy(plt_ras_sweep* h);
int plt_ras_sweep_set_level_rows(plt_ras_sweep* h, int level, const int64_t* value_rows, int64_t n_value,
                                 const int64_t* grad_rows, int64_t n_grad);
int plt_ras_sweep_set_fine(plt_ras_sweep* h, int level, int64_t n_domains, int m, const int64_t* idx, const int32_t* cnt,
                           const double* factor, const double* q_top, const int64_t* inner_glob,
                           const int64_t* inner_loc, int64_t n_inner);
int plt_ras_sweep_set_coarse(plt_ras_sweep* h, int m, const int64_t* idx, const double* inverse, const double* q_top,
                             const double* a_top, const double* p_top_inv);
int plt_ras_sweep_add_transfer(plt_ras_sweep* h, int src_level, int trg_level, int kind, plt_eval* ev);
int plt_ras_sweep_set_poly(plt_ras_sweep* h, const double* p_mono, const double* p_orth, const double* a_p);
int plt_ras_sweep_apply(plt_ras_sweep* h, const double* v, double* out, void* stream);
int64_t plt_ras_sweep_launch_count(plt_ras_sweep* h);
const char* plt_ras_sweep_last_error(plt_ras_sweep* h);

/* The data points on which interpolation::ResidualEvaluator measures the residual exactly
 * (include/polatory/interpolation/residual_evaluator.hpp:123-136): out[0..n) = iota shuffled by a default-seeded
 * std::mt19937 (std::shuffle), then std::partition'ed so that points whose `block` values are not all zero come
 * first.  Host pointers. */
int plt_residual_sample_indices(const double* values, int64_t n, int block, int64_t* out);

/* Library/ABI version and a device probe (returns PLT_ERR_CUDA without a usable GPU). */
int plt_version(void);
int plt_device_check(void);

#ifdef __cplusplus
}
#endif
#endif /* POLATORY_B200_H_ */
