// Host-side construction of the per-RBF constants consumed by rbf.cuh.
// Mirrors the parameter handling of include/polatory/rbf/*.hpp (reference).
#pragma once

#include <cmath>
#include <limits>
#include <stdexcept>
#include <string>

#include "../../include/polatory_b200.h"
#include "rbf.cuh"

namespace plt {

struct RbfHost {
  RbfConst k{};
  bool spheroidal = false;       // needs the direct + fast split when part == FULL
  double support_radius = std::numeric_limits<double>::infinity();  // isotropic
  bool has_hessian = true;
};

inline RbfHost make_rbf_const(int rbf_id, int part, const double* params, int n_params) {
  RbfHost h;
  double p0, p1;
  const bool polyharmonic = rbf_id >= PLT_RBF_BH3 && rbf_id <= PLT_RBF_TH2;
  if (polyharmonic) {
    // polyharmonic_odd.hpp:87-99 / polyharmonic_even.hpp:94-106: {} -> {1, 0}; {s} -> {s, 0}.
    if (n_params < 0 || n_params > 2) throw std::invalid_argument("params.size() must be 2");
    p0 = n_params >= 1 ? params[0] : 1.0;
    p1 = n_params >= 2 ? params[1] : 0.0;
  } else {
    if (n_params != 2) throw std::invalid_argument("params.size() must be 2");  // rbf_base.hpp:81-83
    p0 = params[0];
    p1 = params[1];
  }
  auto& c = h.k.c;
  for (double& v : c) v = 0.0;
  h.k.n = 0;
  if (part != PLT_PART_FULL && !(rbf_id >= PLT_RBF_SP3 && rbf_id <= PLT_RBF_SP9))
    throw std::invalid_argument("only spheroidal RBFs have direct/fast parts");

  auto imq = [&](int n, double B, double C, double D, double E) {
    h.k.n = n;
    c[0] = p0 * B;
    c[1] = p1;
    c[2] = C;
    c[3] = p0 * D / (p1 * p1);
    c[4] = static_cast<double>(n + 2);
    c[5] = E * p1 * p1;
  };

  switch (rbf_id) {
    case PLT_RBF_BH3: h.k.family = FAM_BH3; c[0] = p0; c[1] = p1 * p1; break;
    case PLT_RBF_TH3: h.k.family = FAM_TH3; c[0] = p0; c[1] = p1 * p1; break;
    case PLT_RBF_BH2: h.k.family = FAM_BH2; c[0] = p0; c[1] = p1 * p1; break;
    case PLT_RBF_TH2: h.k.family = FAM_TH2; c[0] = p0; c[1] = p1 * p1; break;
    case PLT_RBF_EXP: h.k.family = FAM_EXP; c[0] = p0; c[1] = p1; break;
    case PLT_RBF_GAU: h.k.family = FAM_GAU; c[0] = p0; c[1] = p1; break;
    // cov_generalized_cauchy{3,5,7}.hpp:24: phi = psill / (1 + a rho^2)^(n/2),
    // g = -a n psill / (...), h = -a (n+2) / (a r^2 + range^2) = -(n+2) / (r^2 + range^2 / a).
    case PLT_RBF_GC3: { double a = 7.0; h.k.family = FAM_IMQ; imq(3, 1.0, a, a * 3.0, 1.0 / a); break; }
    case PLT_RBF_GC5: { double a = 2.4822022531844965; h.k.family = FAM_IMQ; imq(5, 1.0, a, a * 5.0, 1.0 / a); break; }
    case PLT_RBF_GC7: { double a = 1.438027308408951; h.k.family = FAM_IMQ; imq(7, 1.0, a, a * 7.0, 1.0 / a); break; }
    case PLT_RBF_GC9: h.k.family = FAM_IMQ; imq(9, 1.0, 1.0, 9.0, 1.0); break;
    case PLT_RBF_SP3: case PLT_RBF_SP5: case PLT_RBF_SP7: case PLT_RBF_SP9: {
      // cov_spheroidal{3,5,7,9}.hpp:27-32.
      double rho0, A, B, C, D, E; int n;
      if (rbf_id == PLT_RBF_SP3) { n = 3; rho0 = 0.18657871684006438; A = 2.009875543958482; B = 0.8734640537108553; C = 7.181510581693163; D = 18.81837403335934; E = 0.1392464703107397; }
      else if (rbf_id == PLT_RBF_SP5) { n = 5; rho0 = 0.2580127411803573; A = 1.6149073288415876; B = 0.8575980168032007; C = 2.5036086535164204; D = 10.735449080535068; E = 0.39942344766841226; }
      else if (rbf_id == PLT_RBF_SP7) { n = 7; rho0 = 0.2944149476843637; A = 1.4859979204216045; B = 0.8494862533016855; C = 1.44208314742683; D = 8.57520866899984; E = 0.6934412913598931; }
      else { n = 9; rho0 = 0.31622776601683794; A = 1.4230249470757708; B = 0.8445585690332554; C = 1.0; D = 7.601027121299299; E = 1.0; }
      imq(n, B, C, D, E);
      c[6] = rho0;
      c[7] = p0 * A;
      c[8] = p0;
      h.spheroidal = true;
      if (part == PLT_PART_FAST) h.k.family = FAM_IMQ;
      else if (part == PLT_PART_DIRECT) { h.k.family = FAM_SPD_DIRECT; h.support_radius = rho0 * p1; }
      else h.k.family = FAM_SPD_FULL;
      break;
    }
    case PLT_RBF_SPH: h.k.family = FAM_SPH; c[0] = p0; c[1] = p1; h.support_radius = p1; h.has_hessian = false; break;
    case PLT_RBF_CUB: h.k.family = FAM_CUB; c[0] = p0; c[1] = p1; h.support_radius = p1; h.has_hessian = false; break;
    default: throw std::invalid_argument("unknown RBF id");
  }
  return h;
}

}  // namespace plt
