// polatory_b200_shim.hpp -- header-only C++ side of the drop-in boundary.
//
// Two layers over the C ABI of polatory_b200.h:
//
//  1. plt::Evaluator, plt::Fgmres -- dependency-free RAII wrappers (std::vector in / out, status codes
//     turned back into the exceptions the reference throws).  Always available; this is what the
//     tests compile.
//
//  2. With -DPOLATORY_B200_WITH_POLATORY (i.e. inside a Polatory build, Eigen and the Polatory
//     headers on the include path): polatory::fmm::B200Evaluator<Dim> /
//     B200SymmetricEvaluator<Dim>, which derive from the reference's abstract bases
//       FmmGenericEvaluatorBase<Dim>           include/polatory/fmm/fmm_evaluator.hpp:17-41
//       FmmGenericSymmetricEvaluatorBase<Dim>  include/polatory/fmm/fmm_symmetric_evaluator.hpp:16-37
//     plus the six factory bodies that replace src/fmm/make_fmm_evaluator.cpp:40-270.
//     INTEGRATION.md shows the three-line change to the reference's src/CMakeLists.txt.
#ifndef POLATORY_B200_SHIM_HPP_
#define POLATORY_B200_SHIM_HPP_

#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "polatory_b200.h"

namespace plt {

// Short names of include/polatory/rbf/make_rbf.hpp:30-45 -> PLT_RBF_* ids; -1 if unknown.
inline int rbf_id_from_short_name(const std::string& name) {
  static const char* const names[] = {"bh3", "th3", "bh2", "th2", "exp", "gau", "gc3", "gc5",
                                      "gc7", "gc9", "sp3", "sp5", "sp7", "sp9", "sph", "cub"};
  for (int i = 0; i < 16; ++i)
    if (name == names[i]) return i;
  return -1;
}

// The reference reports every failure on this path as an exception; the status code says which.
inline void throw_status(int status, const char* message) {
  const std::string msg = message && *message ? message : "polatory_b200: status " + std::to_string(status);
  if (status == PLT_ERR_INVALID) throw std::invalid_argument(msg);
  throw std::runtime_error(msg);  // accuracy unattainable, unsupported Hessian, CUDA failure
}

class Evaluator {
 public:
  // kind: PLT_KIND_*; symmetric: sources == targets (the matvec); params / aniso / bbox as in
  // plt_eval_create.  aniso may be empty (identity).
  Evaluator(int kind, bool symmetric, int dim, int rbf_id, const std::vector<double>& params,
            const std::vector<double>& aniso, const double* bbox_min, const double* bbox_max,
            int rbf_part = PLT_PART_FULL)
      : dim_(dim),
        km_((kind == PLT_KIND_F || kind == PLT_KIND_H) ? dim : 1),
        kn_((kind == PLT_KIND_FT || kind == PLT_KIND_H) ? dim : 1),
        symmetric_(symmetric) {
    const int st = plt_eval_create(kind, symmetric ? 1 : 0, dim, rbf_id, rbf_part, params.data(),
                                   static_cast<int>(params.size()), aniso.empty() ? nullptr : aniso.data(),
                                   bbox_min, bbox_max, &h_);
    if (st != PLT_OK) throw_status(st, plt_last_error(nullptr));
  }
  ~Evaluator() { plt_eval_destroy(h_); }
  Evaluator(const Evaluator&) = delete;
  Evaluator& operator=(const Evaluator&) = delete;

  int km() const { return km_; }
  int kn() const { return kn_; }
  plt_eval* handle() const { return h_; }

  void set_accuracy(double accuracy) { check(plt_eval_set_accuracy(h_, accuracy)); }
  // points: row-major n x dim, original (un-transformed) coordinates.
  void set_source_points(const double* points, int64_t n) {
    check(plt_eval_set_source_points(h_, points, n));
    n_src_ = n;
  }
  void set_target_points(const double* points, int64_t n) {
    check(plt_eval_set_target_points(h_, points, n));
    n_trg_ = n;
  }
  void set_points(const double* points, int64_t n) {
    check(plt_eval_set_points(h_, points, n));
    n_src_ = n_trg_ = n;
  }
  void set_weights(const double* weights, int64_t len) { check(plt_eval_set_weights(h_, weights, len)); }
  int64_t result_size() const { return kn_ * (symmetric_ ? n_src_ : n_trg_); }
  void evaluate(double* out, int64_t len) const { check(plt_eval_evaluate(h_, out, len)); }
  std::vector<double> evaluate() const {
    std::vector<double> out(static_cast<size_t>(result_size()));
    evaluate(out.data(), static_cast<int64_t>(out.size()));
    return out;
  }
  // set_target_points(points) + evaluate(out) as one call (interpolation/evaluator.hpp:83-87); host buffers are
  // streamed in slabs through a copy-in / evaluate / copy-out pipeline.
  void evaluate_points(const double* points, int64_t n, double* out, int64_t len) {
    check(plt_eval_evaluate_points(h_, points, n, out, len));
    n_trg_ = n;
  }

 private:
  void check(int status) const {
    if (status != PLT_OK) throw_status(status, plt_last_error(h_));
  }
  plt_eval* h_ = nullptr;
  int dim_, km_, kn_;
  bool symmetric_;
  int64_t n_src_ = 0, n_trg_ = 0;
};

// Device-resident flexible GMRES with the method names of krylov::Fgmres / GmresBase
// (include/polatory/krylov/gmres_base.hpp:11-91, fgmres.hpp:11-26).  The operator and the right preconditioner
// are callables  int(const double* x_dev, double* y_dev)  on DEVICE pointers (return 0 on success); rhs / x0 /
// solutions may be host or device pointers.
class Fgmres {
 public:
  using LinOp = int (*)(void* ctx, const double* x_dev, double* y_dev);

  Fgmres(LinOp op, void* op_ctx, const double* rhs, int64_t n, int max_iter) : rhs_(rhs), n_(n), max_iter_(max_iter) {
    int st = plt_fgmres_create(n, max_iter, &h_);
    if (st != PLT_OK) throw_status(st, plt_fgmres_last_error(nullptr));
    check(plt_fgmres_set_operator(h_, op, op_ctx));
  }
  ~Fgmres() { plt_fgmres_destroy(h_); }
  Fgmres(const Fgmres&) = delete;
  Fgmres& operator=(const Fgmres&) = delete;

  void set_initial_solution(const double* x0) { x0_ = x0; }
  void set_right_preconditioner(LinOp pc, void* ctx) { check(plt_fgmres_set_right_preconditioner(h_, pc, ctx)); }
  [[noreturn]] void set_left_preconditioner(LinOp, void*) {
    throw std::runtime_error("set_left_preconditioner is not supported");  // fgmres.hpp:15-17
  }
  void set_allreduce(plt_allreduce_fn fn, void* ctx) { check(plt_fgmres_set_allreduce(h_, fn, ctx)); }
  void setup() { check(plt_fgmres_setup(h_, rhs_, x0_)); }
  void iterate_process() { check(plt_fgmres_iterate(h_)); }
  void solution_vector(double* x) { check(plt_fgmres_solution(h_, x)); }
  std::vector<double> solution_vector() {
    std::vector<double> x(static_cast<size_t>(n_));
    solution_vector(x.data());
    return x;
  }
  int iteration_count() const { return status().iter; }
  int max_iterations() const { return max_iter_; }
  double absolute_residual() const { return status().abs; }
  double relative_residual() const { return status().rel; }

 private:
  struct Status {
    int iter;
    double abs, rel;
  };
  Status status() const {
    Status s{};
    check(plt_fgmres_status(h_, &s.iter, &s.abs, &s.rel));
    return s;
  }
  void check(int st) const {
    if (st != PLT_OK) throw_status(st, plt_fgmres_last_error(h_));
  }
  plt_fgmres* h_ = nullptr;
  const double* rhs_;
  const double* x0_ = nullptr;
  int64_t n_;
  int max_iter_;
};

// preconditioner::RasPreconditioner::operator() (include/polatory/preconditioner/ras_preconditioner.hpp:183-246) over the
// device tables of a set-up made through plt_ras_* / plt_chol_* / plt_eval_gram_* (INTEGRATION.md section 5): usable
// directly as the right preconditioner of plt::Fgmres (`RasSweep::linop`, ctx = the object).
class RasSweep {
 public:
  RasSweep(int64_t m_rows, int l, int n_levels) {
    int st = plt_ras_sweep_create(m_rows, l, n_levels, &h_);
    if (st != PLT_OK) throw_status(st, "plt_ras_sweep_create");
  }
  ~RasSweep() { plt_ras_sweep_destroy(h_); }
  RasSweep(const RasSweep&) = delete;
  RasSweep& operator=(const RasSweep&) = delete;

  void set_level_rows(int level, const int64_t* value_rows, int64_t n_value, const int64_t* grad_rows, int64_t n_grad) {
    check(plt_ras_sweep_set_level_rows(h_, level, value_rows, n_value, grad_rows, n_grad));
  }
  void set_fine(int level, int64_t n_domains, int m, const int64_t* idx, const int32_t* cnt, const double* factor,
                const double* q_top, const int64_t* inner_glob, const int64_t* inner_loc, int64_t n_inner) {
    check(plt_ras_sweep_set_fine(h_, level, n_domains, m, idx, cnt, factor, q_top, inner_glob, inner_loc, n_inner));
  }
  void set_coarse(int m, const int64_t* idx, const double* inverse, const double* q_top, const double* a_top,
                  const double* p_top_inv) {
    check(plt_ras_sweep_set_coarse(h_, m, idx, inverse, q_top, a_top, p_top_inv));
  }
  void add_transfer(int src_level, int trg_level, int kind, plt_eval* ev) {
    check(plt_ras_sweep_add_transfer(h_, src_level, trg_level, kind, ev));
  }
  void set_poly(const double* p_mono, const double* p_orth, const double* a_p) {
    check(plt_ras_sweep_set_poly(h_, p_mono, p_orth, a_p));
  }
  void operator()(const double* v_dev, double* out_dev, void* stream = nullptr) {
    check(plt_ras_sweep_apply(h_, v_dev, out_dev, stream));
  }
  static int linop(void* ctx, const double* x_dev, double* y_dev) {
    return plt_ras_sweep_apply(static_cast<RasSweep*>(ctx)->h_, x_dev, y_dev, nullptr);
  }

 private:
  void check(int st) const {
    if (st != PLT_OK) throw_status(st, plt_ras_sweep_last_error(h_));
  }
  plt_ras_sweep* h_ = nullptr;
};

}  // namespace plt

#ifdef POLATORY_B200_WITH_POLATORY
// ---------------------------------------------------------------------------------------------
// Adapter onto the reference's own abstract bases (compiled only inside a Polatory build).
// ---------------------------------------------------------------------------------------------
#include <polatory/fmm/fmm_evaluator.hpp>
#include <polatory/fmm/fmm_symmetric_evaluator.hpp>

namespace polatory::fmm {

namespace b200_detail {
template <int Dim>
plt::Evaluator* create(int kind, bool symmetric, const rbf::Rbf<Dim>& rbf, const geometry::Bbox<Dim>& bbox) {
  const int id = plt::rbf_id_from_short_name(rbf.short_name());
  if (id < 0) throw std::runtime_error("not implemented");  // src/fmm/make_fmm_evaluator.cpp:68
  // anisotropy(): Eigen::Matrix<double, Dim, Dim, RowMajor> (include/polatory/types.hpp) -> row-major copy
  const auto& a = rbf.anisotropy();
  std::vector<double> aniso(Dim * Dim);
  for (int i = 0; i < Dim; ++i)
    for (int j = 0; j < Dim; ++j) aniso[i * Dim + j] = a(i, j);
  double lo[Dim], hi[Dim];
  for (int i = 0; i < Dim; ++i) {
    lo[i] = bbox.min()(i);
    hi[i] = bbox.max()(i);
  }
  return new plt::Evaluator(kind, symmetric, Dim, id, rbf.parameters(), aniso, lo, hi);
}
}  // namespace b200_detail

template <int Dim>
class B200Evaluator final : public FmmGenericEvaluatorBase<Dim> {
  using Points = geometry::Points<Dim>;

 public:
  B200Evaluator(int kind, const rbf::Rbf<Dim>& rbf, const geometry::Bbox<Dim>& bbox)
      : impl_(b200_detail::create<Dim>(kind, false, rbf, bbox)) {}

  VecX evaluate() const override {
    VecX y(impl_->result_size());
    impl_->evaluate(y.data(), y.size());
    return y;
  }
  void set_accuracy(double accuracy) override { impl_->set_accuracy(accuracy); }
  // geometry::Points<Dim> is row-major N x Dim and contiguous (include/polatory/geometry/point3d.hpp:29-30)
  void set_source_points(const Points& points) override { impl_->set_source_points(points.data(), points.rows()); }
  void set_target_points(const Points& points) override { impl_->set_target_points(points.data(), points.rows()); }
  void set_weights(const Eigen::Ref<const VecX>& weights) override {
    impl_->set_weights(weights.data(), weights.size());
  }

 private:
  std::unique_ptr<plt::Evaluator> impl_;
};

template <int Dim>
class B200SymmetricEvaluator final : public FmmGenericSymmetricEvaluatorBase<Dim> {
  using Points = geometry::Points<Dim>;

 public:
  B200SymmetricEvaluator(int kind, const rbf::Rbf<Dim>& rbf, const geometry::Bbox<Dim>& bbox)
      : impl_(b200_detail::create<Dim>(kind, true, rbf, bbox)) {}

  VecX evaluate() const override {
    VecX y(impl_->result_size());
    impl_->evaluate(y.data(), y.size());
    return y;
  }
  void set_accuracy(double accuracy) override { impl_->set_accuracy(accuracy); }
  void set_points(const Points& points) override { impl_->set_points(points.data(), points.rows()); }
  void set_weights(const Eigen::Ref<const VecX>& weights) override {
    impl_->set_weights(weights.data(), weights.size());
  }

 private:
  std::unique_ptr<plt::Evaluator> impl_;
};

// The six factories (include/polatory/fmm/fmm_evaluator.hpp:92-106,
// include/polatory/fmm/fmm_symmetric_evaluator.hpp:80-86).  A TU that defines
// POLATORY_B200_DEFINE_FACTORIES replaces src/fmm/make_fmm_evaluator.cpp and
// src/fmm/make_fmm_symmetric_evaluator.cpp (and makes the 24 src/fmm/impl/*.cpp TUs unnecessary).
#ifdef POLATORY_B200_DEFINE_FACTORIES
#define POLATORY_B200_FACTORY(NAME, KIND)                                                        \
  template <int Dim>                                                                             \
  FmmGenericEvaluatorPtr<Dim> NAME(const rbf::Rbf<Dim>& rbf, const geometry::Bbox<Dim>& bbox) {  \
    return std::make_unique<B200Evaluator<Dim>>(KIND, rbf, bbox);                                \
  }                                                                                              \
  template FmmGenericEvaluatorPtr<1> NAME<1>(const rbf::Rbf<1>&, const geometry::Bbox<1>&);      \
  template FmmGenericEvaluatorPtr<2> NAME<2>(const rbf::Rbf<2>&, const geometry::Bbox<2>&);      \
  template FmmGenericEvaluatorPtr<3> NAME<3>(const rbf::Rbf<3>&, const geometry::Bbox<3>&);
POLATORY_B200_FACTORY(make_fmm_evaluator, PLT_KIND_K)
POLATORY_B200_FACTORY(make_fmm_gradient_evaluator, PLT_KIND_F)
POLATORY_B200_FACTORY(make_fmm_gradient_transpose_evaluator, PLT_KIND_FT)
POLATORY_B200_FACTORY(make_fmm_hessian_evaluator, PLT_KIND_H)
#undef POLATORY_B200_FACTORY
#define POLATORY_B200_SYM_FACTORY(NAME, KIND)                                                            \
  template <int Dim>                                                                                     \
  FmmGenericSymmetricEvaluatorPtr<Dim> NAME(const rbf::Rbf<Dim>& rbf, const geometry::Bbox<Dim>& bbox) { \
    return std::make_unique<B200SymmetricEvaluator<Dim>>(KIND, rbf, bbox);                               \
  }                                                                                                      \
  template FmmGenericSymmetricEvaluatorPtr<1> NAME<1>(const rbf::Rbf<1>&, const geometry::Bbox<1>&);     \
  template FmmGenericSymmetricEvaluatorPtr<2> NAME<2>(const rbf::Rbf<2>&, const geometry::Bbox<2>&);     \
  template FmmGenericSymmetricEvaluatorPtr<3> NAME<3>(const rbf::Rbf<3>&, const geometry::Bbox<3>&);
POLATORY_B200_SYM_FACTORY(make_fmm_symmetric_evaluator, PLT_KIND_K)
POLATORY_B200_SYM_FACTORY(make_fmm_hessian_symmetric_evaluator, PLT_KIND_H)
#undef POLATORY_B200_SYM_FACTORY
#endif  // POLATORY_B200_DEFINE_FACTORIES

}  // namespace polatory::fmm
#endif  // POLATORY_B200_WITH_POLATORY

#endif  // POLATORY_B200_SHIM_HPP_
