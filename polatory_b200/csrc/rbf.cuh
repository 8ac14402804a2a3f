// Device-side RBF formulas for the near-field (P2P) and M2L-operator kernels.
//
// Every RBF of include/polatory/rbf/*.hpp (reference) is reduced to three radial
// scalars of the isotropic (anisotropy-transformed) difference d, r2 = |d|^2:
//
//   phi(d)      = PHI(r2)
//   grad phi    = G(r2) * d
//   hess phi    = G(r2) * I + GH(r2) * d d^T          (GH = G * h of SURVEY.md 8a)
//
// The anisotropy matrix A is folded into the weights before and into the outputs after
// the pair loop (w' = A w, out = A^T v), so the pair loop only sees isotropic blocks:
//
//   K  : v   +=  PHI * w                     include/polatory/fmm/kernel.hpp:43-52
//   F  : v   += -G * (d . w')                include/polatory/fmm/gradient_kernel.hpp:45-60
//   FT : v_b +=  G * d_b * w                 include/polatory/fmm/gradient_transpose_kernel.hpp:47-62
//   H  : v_b += -(G * w'_b + GH * (d . w') d_b)   include/polatory/fmm/hessian_kernel.hpp:46-65
#pragma once

#include <cuda_runtime.h>
#include <math.h>

namespace plt {

enum Family : int {
  FAM_BH3 = 0,    // polyharmonic_odd.hpp K=1
  FAM_TH3,        // polyharmonic_odd.hpp K=3
  FAM_BH2,        // polyharmonic_even.hpp K=2
  FAM_TH2,        // polyharmonic_even.hpp K=4
  FAM_EXP,        // cov_exponential.hpp
  FAM_GAU,        // cov_gaussian.hpp
  FAM_IMQ,        // cov_generalized_cauchy{3,5,7,9}.hpp and spheroidal fast parts
  FAM_SPD_FULL,   // cov_spheroidal{3,5,7,9}.hpp kFull   (only used by brute force / P2P)
  FAM_SPD_DIRECT, // cov_spheroidal{3,5,7,9}.hpp kDirectPart
  FAM_SPH,        // cov_spherical.hpp
  FAM_CUB,        // cov_cubic.hpp
  FAM_COUNT
};

enum Kind : int { KIND_K = 0, KIND_F = 1, KIND_FT = 2, KIND_H = 3 };

// What the pair loop needs from the radial functions.
enum Need : int { NEED_PHI = 0, NEED_G = 1, NEED_G_GH = 2 };

// Constants of one RBF instance, precomputed on the host (rbf_host.cpp).
struct RbfConst {
  int family;
  int n;        // IMQ exponent numerator: phi ~ t^(-n/2)
  double c[10];
  // BH3/TH3/BH2/TH2: c0 = slope, c1 = c*c
  // EXP/GAU/SPH/CUB: c0 = psill, c1 = range
  // IMQ (+SPD): c0 = psill*B, c1 = range, c2 = C, c3 = psill*D/range^2, c4 = n+2,
  //             c5 = E*range^2, c6 = rho0, c7 = psill*A (lin), c8 = psill
};

template <int KIND, int DIM>
struct KindTraits {
  static constexpr int km = (KIND == KIND_F || KIND == KIND_H) ? DIM : 1;
  static constexpr int kn = (KIND == KIND_FT || KIND == KIND_H) ? DIM : 1;
  static constexpr int need = KIND == KIND_K ? NEED_PHI : (KIND == KIND_H ? NEED_G_GH : NEED_G);
};

__device__ __forceinline__ double ipow_odd(double q, double q2, int n) {
  // q^n for odd n >= 1, q2 = q*q
  double r = q;
  for (int i = 1; i < n; i += 2) r *= q2;
  return r;
}

// 1/sqrt(x) for normal x > 0: MUFU.RSQ64H seed (relative error ~2^-21) + one third-order
// correction y (1 + e/2 + 3 e^2 / 8), e = 1 - x y^2  -> error ~e^3 < 2^-60, i.e. correctly
// rounded to within the last ulp or two.  Branch-free, 5 FP64-pipe instructions; CUDA's
// rsqrt(double) costs ~10 plus an integer special-case prologue, which made the bh3 pair loop
// issue-bound at 37 % FP64-pipe utilisation (profiles/r01_f_p2p_full.md).
__device__ __forceinline__ double rsqrt_seed(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  return y;
}
__device__ __forceinline__ double rsqrt_fast(double x) {
  const double y = rsqrt_seed(x);
  const double xy = x * y;
  const double e = fma(-xy, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  return fma(y * e, p, y);
}
// sqrt(x) for x >= 0 (returns 0 at 0): same seed, corrected as xy (1 + e p).
__device__ __forceinline__ double sqrt_fast(double x) {
  const double xc = x + 1e-300;  // keeps the seed finite at x == 0; sqrt(1e-300) = 1e-150 ~ 0
  const double y = rsqrt_seed(xc);
  const double xy = xc * y;
  const double e = fma(-xy, y, 1.0);
  const double p = fma(0.375, e, 0.5);
  return fma(xy * e, p, xy);
}

// Radial scalars.  NEED selects which outputs are computed (others untouched).
template <int FAM, int NEED>
__device__ __forceinline__ void rbf_radial(const RbfConst& k, double r2, double& phi, double& g,
                                           double& gh) {
  if constexpr (FAM == FAM_BH3) {
    // phi = -s rho, g = -s / rho, gh = +s / rho^3; zero gradient/Hessian at rho == 0
    // (polyharmonic_odd.hpp:32-67, K = 1, kSign = -1).
    double rho2 = r2 + k.c[1];
    if constexpr (NEED == NEED_PHI) {
      phi = -k.c[0] * sqrt_fast(rho2);
    } else {
      double inv = rho2 > 0.0 ? rsqrt_fast(rho2 + 1e-300) : 0.0;
      g = -k.c[0] * inv;
      if constexpr (NEED == NEED_G_GH) gh = k.c[0] * inv * inv * inv;
    }
  } else if constexpr (FAM == FAM_TH3) {
    // phi = s rho^3, g = 3 s rho, gh = 3 s / rho (K = 3, kSign = +1).
    double rho2 = r2 + k.c[1];
    double inv = 0.0, rho;
    if constexpr (NEED == NEED_PHI) {
      rho = sqrt_fast(rho2);
    } else {
      inv = rho2 > 0.0 ? rsqrt_fast(rho2 + 1e-300) : 0.0;
      rho = rho2 * inv;
    }
    if constexpr (NEED == NEED_PHI) phi = k.c[0] * rho2 * rho;
    if constexpr (NEED >= NEED_G) g = 3.0 * k.c[0] * rho;
    if constexpr (NEED == NEED_G_GH) gh = 3.0 * k.c[0] * inv;
  } else if constexpr (FAM == FAM_BH2) {
    // phi = s rho^2 ln rho, g = s (1 + 2 ln rho), gh = 2 s / rho^2
    // (polyharmonic_even.hpp:33-73, K = 2, kSign = +1); all zero at rho == 0.
    double rho2 = r2 + k.c[1];
    bool nz = rho2 > 0.0;
    double l = nz ? 0.5 * log(rho2) : 0.0;
    if constexpr (NEED == NEED_PHI) phi = k.c[0] * rho2 * l;
    if constexpr (NEED >= NEED_G) g = nz ? k.c[0] * (1.0 + 2.0 * l) : 0.0;
    if constexpr (NEED == NEED_G_GH) gh = nz ? 2.0 * k.c[0] / rho2 : 0.0;
  } else if constexpr (FAM == FAM_TH2) {
    // phi = -s rho^4 ln rho, g = -s rho^2 (1 + 4 ln rho), gh = -s (6 + 8 ln rho) (K = 4).
    double rho2 = r2 + k.c[1];
    bool nz = rho2 > 0.0;
    double l = nz ? 0.5 * log(rho2) : 0.0;
    if constexpr (NEED == NEED_PHI) phi = -k.c[0] * rho2 * rho2 * l;
    if constexpr (NEED >= NEED_G) g = -k.c[0] * rho2 * (1.0 + 4.0 * l);
    if constexpr (NEED == NEED_G_GH) gh = nz ? -k.c[0] * (6.0 + 8.0 * l) : 0.0;
  } else if constexpr (FAM == FAM_EXP) {
    // cov_exponential.hpp:33-61.  r == 0 gives inf/NaN for g, gh exactly as the reference.
    double r = sqrt(r2);
    double e = k.c[0] * exp(-3.0 * (r / k.c[1]));
    if constexpr (NEED == NEED_PHI) phi = e;
    if constexpr (NEED >= NEED_G) g = -3.0 * e / (k.c[1] * r);
    if constexpr (NEED == NEED_G_GH) gh = -g * (1.0 / r2 + 3.0 / (k.c[1] * r));
  } else if constexpr (FAM == FAM_GAU) {
    // cov_gaussian.hpp:33-60.
    double irr = 1.0 / (k.c[1] * k.c[1]);
    double e = k.c[0] * exp(-3.0 * r2 * irr);
    if constexpr (NEED == NEED_PHI) phi = e;
    if constexpr (NEED >= NEED_G) g = -6.0 * e * irr;
    if constexpr (NEED == NEED_G_GH) gh = -g * 6.0 * irr;
  } else if constexpr (FAM == FAM_IMQ || FAM == FAM_SPD_FULL || FAM == FAM_SPD_DIRECT) {
    // imq(t) = psill B t^(-n/2), t = 1 + C rho^2, rho = r / range
    // (cov_generalized_cauchy3.hpp:35-63, cov_spheroidal3.hpp:44-108).
    double irr = 1.0 / (k.c[1] * k.c[1]);
    double t = 1.0 + k.c[2] * (r2 * irr);
    double q = rsqrt(t), q2 = q * q;
    double qn = ipow_odd(q, q2, k.n);
    double p_i = 0.0, g_i = 0.0, gh_i = 0.0;
    if constexpr (NEED == NEED_PHI) p_i = k.c[0] * qn;
    if constexpr (NEED >= NEED_G) g_i = -k.c[3] * qn * q2;
    if constexpr (NEED == NEED_G_GH) gh_i = -g_i * k.c[4] / (r2 + k.c[5]);
    if constexpr (FAM == FAM_IMQ) {
      if constexpr (NEED == NEED_PHI) phi = p_i;
      if constexpr (NEED >= NEED_G) g = g_i;
      if constexpr (NEED == NEED_G_GH) gh = gh_i;
    } else {
      // lin = psill (1 - A rho) inside rho < rho0.
      double r = sqrt(r2);
      double rho = r / k.c[1];
      bool in = rho < k.c[6];
      double p_l = 0.0, g_l = 0.0, gh_l = 0.0;
      if constexpr (NEED == NEED_PHI) p_l = k.c[8] - k.c[7] * rho;
      if constexpr (NEED >= NEED_G) g_l = -k.c[7] / (r * k.c[1]);
      if constexpr (NEED == NEED_G_GH) gh_l = -g_l / r2;
      if constexpr (FAM == FAM_SPD_FULL) {
        if constexpr (NEED == NEED_PHI) phi = in ? p_l : p_i;
        if constexpr (NEED >= NEED_G) g = in ? g_l : g_i;
        if constexpr (NEED == NEED_G_GH) gh = in ? gh_l : gh_i;
      } else {
        if constexpr (NEED == NEED_PHI) phi = in ? p_l - p_i : 0.0;
        if constexpr (NEED >= NEED_G) g = in ? g_l - g_i : 0.0;
        if constexpr (NEED == NEED_G_GH) gh = in ? gh_l - gh_i : 0.0;
      }
    }
  } else if constexpr (FAM == FAM_SPH) {
    // cov_spherical.hpp:34-51 (Hessian throws in the reference; rejected on the host).
    double r = sqrt(r2);
    double rho = r / k.c[1];
    bool in = r < k.c[1];
    if constexpr (NEED == NEED_PHI) phi = in ? k.c[0] * (1.0 + rho * (-1.5 + 0.5 * rho * rho)) : 0.0;
    if constexpr (NEED >= NEED_G) g = in ? k.c[0] * (-1.5 / rho + 1.5 * rho) / (k.c[1] * k.c[1]) : 0.0;
    if constexpr (NEED == NEED_G_GH) gh = 0.0;
  } else if constexpr (FAM == FAM_CUB) {
    // cov_cubic.hpp:34-56.
    double r = sqrt(r2);
    double rho = r / k.c[1];
    double rho2 = rho * rho;
    bool in = r < k.c[1];
    if constexpr (NEED == NEED_PHI)
      phi = in ? k.c[0] * (1.0 + rho2 * (-7.0 + rho * (8.75 + rho2 * (-3.5 + 0.75 * rho2)))) : 0.0;
    if constexpr (NEED >= NEED_G)
      g = in ? k.c[0] * (-14.0 + rho * (26.25 + rho2 * (-17.5 + 5.25 * rho2))) / (k.c[1] * k.c[1]) : 0.0;
    if constexpr (NEED == NEED_G_GH) gh = 0.0;
  }
}

// One pair: accumulate v[kn] += block(d) * w[km] on the isotropic difference d[DIM].
template <int FAM, int KIND, int DIM>
__device__ __forceinline__ void pair_accumulate(const RbfConst& k, const double (&d)[DIM],
                                                const double* __restrict__ w, double* v) {
  double r2 = 0.0;
#pragma unroll
  for (int a = 0; a < DIM; ++a) r2 = fma(d[a], d[a], r2);
  double phi = 0.0, g = 0.0, gh = 0.0;
  rbf_radial<FAM, KindTraits<KIND, DIM>::need>(k, r2, phi, g, gh);
  if constexpr (KIND == KIND_K) {
    v[0] = fma(phi, w[0], v[0]);
  } else if constexpr (KIND == KIND_F) {
    double dw = 0.0;
#pragma unroll
    for (int a = 0; a < DIM; ++a) dw = fma(d[a], w[a], dw);
    v[0] = fma(-g, dw, v[0]);
  } else if constexpr (KIND == KIND_FT) {
    double gw = g * w[0];
#pragma unroll
    for (int b = 0; b < DIM; ++b) v[b] = fma(gw, d[b], v[b]);
  } else {
    double dw = 0.0;
#pragma unroll
    for (int a = 0; a < DIM; ++a) dw = fma(d[a], w[a], dw);
    double s = gh * dw;
#pragma unroll
    for (int b = 0; b < DIM; ++b) v[b] -= fma(g, w[b], s * d[b]);
  }
}

// The kn x km block itself (used to tabulate M2L operators): blk[b*km + a].
template <int FAM, int KIND, int DIM>
__device__ __forceinline__ void kernel_block(const RbfConst& k, const double (&d)[DIM], double* blk) {
  double r2 = 0.0;
#pragma unroll
  for (int a = 0; a < DIM; ++a) r2 = fma(d[a], d[a], r2);
  double phi = 0.0, g = 0.0, gh = 0.0;
  rbf_radial<FAM, KindTraits<KIND, DIM>::need>(k, r2, phi, g, gh);
  if constexpr (KIND == KIND_K) {
    blk[0] = phi;
  } else if constexpr (KIND == KIND_F) {
#pragma unroll
    for (int a = 0; a < DIM; ++a) blk[a] = -g * d[a];
  } else if constexpr (KIND == KIND_FT) {
#pragma unroll
    for (int b = 0; b < DIM; ++b) blk[b] = g * d[b];
  } else {
#pragma unroll
    for (int b = 0; b < DIM; ++b)
#pragma unroll
      for (int a = 0; a < DIM; ++a) blk[b * DIM + a] = -((a == b ? g : 0.0) + gh * d[b] * d[a]);
  }
}

}  // namespace plt
