// Host-side tables of the equispaced ("modified uniform") interpolator.
//
// The reference instantiates scalfmm::interpolation::interpolator<double, Dim, Kernel,
// options::modified_uniform_>(kernel, order, tree_height, box_width, d)
// (src/fmm/fmm_evaluator.hpp:57-58,249-250).  That class lives in the un-vendored
// ScalFMM3 fork; what is restated here is its published construction:
//   * `order` equispaced nodes per axis on the cell, end points included,
//     t_i = -1 + 2 i / (order - 1);
//   * d = kClassic (-1): the polynomial (Lagrange) interpolant; d in 0..order-1: the
//     Floater-Hormann rational interpolant of degree d
//     (src/fmm/interpolator_configuration.hpp:9-19), both in barycentric form
//     S_i(t) = (beta_i / (t - t_i)) / sum_j (beta_j / (t - t_j));
//   * M2M / L2L = interpolation of the child's nodes by the parent's basis, per axis;
//   * M2L on the (2*order-1)^D difference lattice, diagonalised by a DFT of that length.
#pragma once

#include <cmath>
#include <vector>

namespace plt {

constexpr int kMaxOrder = 20;

struct InterpTables {
  int order = 0;
  int d = -1;
  int nf = 0;  // DFT length 2*order - 1
  std::vector<double> beta;   // [order] barycentric weights
  std::vector<double> child;  // [2][order][order]: child[s][m][n] = S_m(parent)(node n of child s)
  std::vector<double> tw;     // [nf][2]: cos, -sin of 2 pi j / nf  (forward twiddles e^{-2 pi i j/nf})
};

inline double interp_node(int i, int order) { return -1.0 + 2.0 * i / (order - 1); }

// Floater & Hormann (2007), eq. (18), equispaced nodes: beta_k = (-1)^(k-d) sum_{i in J_k} C(d, k-i),
// J_k = {i : max(0, k-d) <= i <= min(k, n-d)}, n = order-1.  d = n gives the polynomial weights.
inline std::vector<double> barycentric_weights(int order, int d) {
  const int n = order - 1;
  if (d < 0 || d > n) d = n;
  auto binom = [](int a, int b) {
    double r = 1.0;
    for (int i = 1; i <= b; ++i) r = r * (a - b + i) / i;
    return r;
  };
  std::vector<double> beta(order);
  for (int k = 0; k <= n; ++k) {
    double s = 0.0;
    for (int i = std::max(0, k - d); i <= std::min(k, n - d); ++i) s += binom(d, k - i);
    beta[k] = (((k - d) % 2 + 2) % 2 == 0 ? 1.0 : -1.0) * s;
  }
  return beta;
}

inline void barycentric_basis(int order, const double* beta, double t, double* s) {
  int hit = -1;
  double sum = 0.0;
  for (int i = 0; i < order; ++i) {
    double dt = t - interp_node(i, order);
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

inline InterpTables make_interp_tables(int order, int d) {
  InterpTables t;
  t.order = order;
  t.d = d;
  t.nf = 2 * order - 1;
  t.beta = barycentric_weights(order, d);
  t.child.assign(2 * order * order, 0.0);
  std::vector<double> s(order);
  for (int side = 0; side < 2; ++side) {
    for (int n = 0; n < order; ++n) {
      // node n of the child, in the parent's [-1, 1] coordinate
      double y = 0.5 * interp_node(n, order) + (side == 0 ? -0.5 : 0.5);
      barycentric_basis(order, t.beta.data(), y, s.data());
      for (int m = 0; m < order; ++m) t.child[(side * order + m) * order + n] = s[m];
    }
  }
  t.tw.resize(2 * t.nf);
  const long double two_pi = 6.283185307179586476925286766559L;
  for (int j = 0; j < t.nf; ++j) {
    long double a = two_pi * j / t.nf;
    t.tw[2 * j] = static_cast<double>(cosl(a));
    t.tw[2 * j + 1] = static_cast<double>(-sinl(a));
  }
  return t;
}

}  // namespace plt
