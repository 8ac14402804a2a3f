/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C + OpenMP) of the reference's
 * RBF fast-multipole evaluation path.  Nothing under polatory_b200/ may link or call this;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do,
 * and only as the checker / the timed CPU baseline.
 *
 * What is restated, with the reference file:line each part follows (paths relative to
 * /root/reference):
 *   - RBF formulas ...................... include/polatory/rbf/ headers (cited per function)
 *   - kernel functors K / F / FT / H .... include/polatory/fmm/kernel.hpp:43-52,
 *                                         gradient_kernel.hpp:45-60,
 *                                         gradient_transpose_kernel.hpp:47-62,
 *                                         hessian_kernel.hpp:46-65
 *   - brute force ....................... src/fmm/full_direct.hpp:7-52
 *   - self interaction .................. src/fmm/fmm_symmetric_evaluator.hpp:163-193
 *   - evaluator driver .................. src/fmm/fmm_evaluator.hpp:77-112,226-270
 *   - tree height / root box ............ src/fmm/utility.hpp:12-33
 *   - interpolator configuration ........ src/fmm/interpolator_configuration.hpp:9-19
 *
 * The FMM passes themselves (P2M, M2M, M2L, L2L, L2P, P2P) live in the reference's
 * un-vendored dependency polatory/ScalFMM3 (floating git tag `polatory`,
 * src/CMakeLists.txt:155-163) which is absent from /root/reference and from this image.
 * They are restated from the published algorithm of that library's uniform-interpolation
 * FMM (Blanchard, Coulaud, Darve: "Fast hierarchical algorithms for generating Gaussian
 * random fields", 2015 -- Lagrange interpolation on equispaced nodes, M2L applied through
 * a circulant embedding + FFT) with the Floater-Hormann rational variant selected by `d`
 * (Floater & Hormann, Numer. Math. 107, 2007).
 *
 * PARITY STATUS: "parity unpinned" against the reference FMM -- the reference's tests hold
 * no golden vectors for this path and the reference cannot be built here.
 * What is pinned: the RBF formulas (finite-difference test of test/rbf/test_rbf.cpp:126-174,
 * restated in tests/test_oracle_rbf.py) and the reference's own acceptance criterion
 * |FMM - direct sum|_inf < accuracy (test/interpolation/test_evaluator.cpp:70-75).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXD 3
enum { K_K = 0, K_F = 1, K_FT = 2, K_H = 3 };
enum { R_BH3, R_TH3, R_BH2, R_TH2, R_EXP, R_GAU, R_GC3, R_GC5, R_GC7, R_GC9, R_SP3, R_SP5, R_SP7, R_SP9,
       R_SPH, R_CUB };
enum { PART_FULL = 0, PART_DIRECT = 1, PART_FAST = 2 };

typedef struct {
  int id, part, dim;
  double p0, p1;      /* parameters() */
  double A[MAXD * MAXD]; /* anisotropy, row-major dim x dim */
} orc_rbf;

/* ------------------------------------------------------------------------------------------
 * RBF formulas.  Each returns the three radial scalars of one RBF on the isotropic
 * difference d:  phi, the gradient coefficient g (grad = g d) and the Hessian factor h
 * (hess = g (I + h d d^T)), evaluated in the reference's operation order.
 * `zero` is set when the reference returns an all-zero gradient/Hessian (rho == 0 cases).
 * ------------------------------------------------------------------------------------------ */
static double sqrt_pow(double x, int n) { /* rbf_base.hpp:131-153 */
  double s = sqrt(x);
  switch (n) {
    case 3: return x * s;
    case 5: return x * x * s;
    case 7: return x * x * x * s;
    case 9: { double x2 = x * x; return x2 * x2 * s; }
    case 11: { double x2 = x * x; return x2 * x2 * x * s; }
  }
  return pow(x, n / 2.0);
}

typedef struct { double rho0, A, B, C, D, E; int n; } spd_const;
static spd_const spd_constants(int id) { /* cov_spheroidal{3,5,7,9}.hpp:27-32 */
  spd_const c;
  switch (id) {
    case R_SP3: c = (spd_const){0.18657871684006438, 2.009875543958482, 0.8734640537108553, 7.181510581693163, 18.81837403335934, 0.1392464703107397, 3}; break;
    case R_SP5: c = (spd_const){0.2580127411803573, 1.6149073288415876, 0.8575980168032007, 2.5036086535164204, 10.735449080535068, 0.39942344766841226, 5}; break;
    case R_SP7: c = (spd_const){0.2944149476843637, 1.4859979204216045, 0.8494862533016855, 1.44208314742683, 8.57520866899984, 0.6934412913598931, 7}; break;
    default:    c = (spd_const){0.31622776601683794, 1.4230249470757708, 0.8445585690332554, 1.0, 7.601027121299299, 1.0, 9}; break;
  }
  return c;
}

static void rbf_radial(const orc_rbf* r, const double* d, double* phi, double* g, double* h, int* zero) {
  const int dim = r->dim;
  double r2 = 0.0;
  for (int a = 0; a < dim; ++a) r2 += d[a] * d[a];
  *zero = 0;
  *phi = *g = *h = 0.0;
  switch (r->id) {
    case R_BH3: case R_TH3: { /* polyharmonic_odd.hpp:32-67 */
      const int K = r->id == R_BH3 ? 1 : 3;
      const double sign = r->id == R_BH3 ? -1.0 : 1.0;
      double slope = r->p0, c = r->p1;
      double rho2 = r2 + c * c, rho = sqrt(rho2);
      *phi = sign * slope * (K == 1 ? rho : rho * rho * rho);
      if (rho == 0.0) { *zero = 1; return; }
      *g = sign * K * slope * (K == 1 ? 1.0 / rho : rho);
      *h = (K - 2) / rho2;
      return;
    }
    case R_BH2: case R_TH2: { /* polyharmonic_even.hpp:33-73 */
      const int K = r->id == R_BH2 ? 2 : 4;
      const double sign = r->id == R_BH2 ? 1.0 : -1.0;
      double slope = r->p0, c = r->p1;
      double rho2 = r2 + c * c, rho = sqrt(rho2);
      if (rho == 0.0) { *zero = 1; return; }
      double powk = K == 2 ? rho * rho : (rho * rho) * (rho * rho);
      double powk2 = K == 2 ? 1.0 : rho * rho;
      *phi = sign * slope * powk * log(rho);
      *g = sign * slope * powk2 * (1.0 + K * log(rho));
      *h = (K - 2.0 + K / (1.0 + K * log(rho))) / rho2;
      return;
    }
    case R_EXP: { /* cov_exponential.hpp:33-61 */
      double psill = r->p0, range = r->p1, rr = sqrt(r2), rho = rr / range;
      *phi = psill * exp(-3.0 * rho);
      *g = -3.0 * psill * exp(-3.0 * rho) / (range * rr);
      *h = -(1.0 / (rr * rr) + 3.0 / (range * rr));
      return;
    }
    case R_GAU: { /* cov_gaussian.hpp:33-60 */
      double psill = r->p0, range = r->p1, rr = sqrt(r2), rho = rr / range;
      *phi = psill * exp(-3.0 * rho * rho);
      *g = -6.0 * psill * exp(-3.0 * rho * rho) / (range * range);
      *h = -6.0 / (range * range);
      return;
    }
    case R_GC3: case R_GC5: case R_GC7: { /* cov_generalized_cauchy{3,5,7}.hpp:24,35-63 */
      const double kA = r->id == R_GC3 ? 7.0 : (r->id == R_GC5 ? 2.4822022531844965 : 1.438027308408951);
      const int n = r->id == R_GC3 ? 3 : (r->id == R_GC5 ? 5 : 7);
      double psill = r->p0, range = r->p1, rr = sqrt(r2), rho = rr / range;
      *phi = psill / sqrt_pow(1.0 + kA * rho * rho, n);
      *g = -kA * (double)n * psill / (sqrt_pow(1.0 + kA * rho * rho, n + 2) * range * range);
      *h = -kA * (double)(n + 2) / (kA * rr * rr + range * range);
      return;
    }
    case R_GC9: { /* cov_generalized_cauchy9.hpp:39-59 */
      double psill = r->p0, range = r->p1, rr = sqrt(r2), rho = rr / range;
      *phi = psill / sqrt_pow(1.0 + rho * rho, 9);
      *g = -9.0 * psill / (sqrt_pow(1.0 + rho * rho, 11) * range * range);
      *h = -11.0 / (rr * rr + range * range);
      return;
    }
    case R_SP3: case R_SP5: case R_SP7: case R_SP9: { /* cov_spheroidal3.hpp:44-108 */
      spd_const c = spd_constants(r->id);
      double psill = r->p0, range = r->p1, rr = sqrt(r2), rho = rr / range;
      double t = r->id == R_SP9 ? 1.0 + rho * rho : 1.0 + c.C * rho * rho;
      double phi_l = psill * (1.0 - c.A * rho);
      double g_l = -psill * c.A / (rr * range);
      double h_l = -1.0 / (rr * rr);
      double phi_i = psill * c.B / sqrt_pow(t, c.n);
      double g_i = -psill * c.D / (sqrt_pow(t, c.n + 2) * range * range);
      double h_i = r->id == R_SP9 ? -(double)(c.n + 2) / (rr * rr + range * range)
                                  : -(double)(c.n + 2) / (rr * rr + c.E * range * range);
      int in = rho < c.rho0;
      if (r->part == PART_FAST) { *phi = phi_i; *g = g_i; *h = h_i; return; }
      if (r->part == PART_FULL) {
        if (in) { *phi = phi_l; *g = g_l; *h = h_l; } else { *phi = phi_i; *g = g_i; *h = h_i; }
        return;
      }
      /* direct part: lin - imq inside, 0 outside.  The difference of two Hessians is
         (g_l - g_i) I + (g_l h_l - g_i h_i) d d^T; encode as g = g_l - g_i, h = that / g. */
      if (!in) { *zero = 1; return; }
      *phi = phi_l - phi_i;
      *g = g_l - g_i;
      *h = (g_l * h_l - g_i * h_i) / (g_l - g_i);
      return;
    }
    case R_SPH: { /* cov_spherical.hpp:34-51 */
      double psill = r->p0, range = r->p1, rr = sqrt(r2), rho = rr / range;
      if (!(rr < range)) { *zero = 1; return; }
      *phi = psill * (1.0 + rho * (-1.5 + 0.5 * rho * rho));
      *g = psill * (-1.5 / rho + 1.5 * rho) / (range * range);
      return;
    }
    case R_CUB: { /* cov_cubic.hpp:34-56 */
      double psill = r->p0, range = r->p1, rr = sqrt(r2), rho = rr / range, rho2 = rho * rho;
      if (!(rr < range)) { *zero = 1; return; }
      *phi = psill * (1.0 + rho2 * (-7.0 + rho * (8.75 + rho2 * (-3.5 + 0.75 * rho2))));
      *g = psill * (-14.0 + rho * (26.25 + rho2 * (-17.5 + 5.25 * rho2))) / (range * range);
      return;
    }
  }
}

static int kind_km(int kind, int dim) { return (kind == K_F || kind == K_H) ? dim : 1; }
static int kind_kn(int kind, int dim) { return (kind == K_FT || kind == K_H) ? dim : 1; }

/* kernel.evaluate(x, y) on transformed positions: k[b*km + a]. */
static void kernel_eval(const orc_rbf* r, int kind, const double* x, const double* y, double* k) {
  const int dim = r->dim;
  double d[MAXD], phi, g, h;
  int zero;
  for (int a = 0; a < dim; ++a) d[a] = x[a] - y[a];
  rbf_radial(r, d, &phi, &g, &h, &zero);
  if (kind == K_K) { k[0] = phi; return; }
  if (kind == K_F || kind == K_FT) {
    /* -(grad_iso * A) or +(grad_iso * A): row vector times matrix */
    for (int a = 0; a < dim; ++a) {
      double s = 0.0;
      if (!zero) for (int c = 0; c < dim; ++c) s += g * d[c] * r->A[c * dim + a];
      k[a] = kind == K_F ? -s : s;
    }
    return;
  }
  /* -(A^T H_iso A) */
  double H[MAXD * MAXD], T[MAXD * MAXD];
  for (int i = 0; i < dim; ++i)
    for (int j = 0; j < dim; ++j) H[i * dim + j] = zero ? 0.0 : g * ((i == j ? 1.0 : 0.0) + h * d[i] * d[j]);
  for (int i = 0; i < dim; ++i)
    for (int j = 0; j < dim; ++j) {
      double s = 0.0;
      for (int c = 0; c < dim; ++c) s += H[i * dim + c] * r->A[c * dim + j];
      T[i * dim + j] = s;
    }
  for (int i = 0; i < dim; ++i)
    for (int j = 0; j < dim; ++j) {
      double s = 0.0;
      for (int c = 0; c < dim; ++c) s += r->A[c * dim + i] * T[c * dim + j];
      k[i * dim + j] = -s;
    }
}

static void transform_points(const orc_rbf* r, const double* pts, int64_t n, double* out) {
  /* geometry/point3d.hpp:36-47: p * A^T */
  const int dim = r->dim;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i)
    for (int a = 0; a < dim; ++a) {
      double s = 0.0;
      for (int b = 0; b < dim; ++b) s += pts[i * dim + b] * r->A[a * dim + b];
      out[i * dim + a] = s;
    }
}

static orc_rbf make_rbf(int id, int part, int dim, const double* params, const double* aniso) {
  orc_rbf r;
  r.id = id; r.part = part; r.dim = dim; r.p0 = params[0]; r.p1 = params[1];
  for (int i = 0; i < dim * dim; ++i) r.A[i] = aniso ? aniso[i] : (i / dim == i % dim ? 1.0 : 0.0);
  return r;
}

/* ------------------------------------------------------------------------------------------
 * Brute force: src/fmm/full_direct.hpp:7-52 (+ self interaction for the symmetric variant).
 * Points in ORIGINAL coordinates.  out: kn per target.
 * ------------------------------------------------------------------------------------------ */
int orc_direct(int rbf_id, int part, int dim, const double* params, const double* aniso, int kind,
               const double* src, int64_t ns, const double* trg, int64_t nt, const double* w, int symmetric,
               double* out) {
  orc_rbf r = make_rbf(rbf_id, part, dim, params, aniso);
  const int km = kind_km(kind, dim), kn = kind_kn(kind, dim);
  double* s = (double*)malloc(sizeof(double) * (ns > 0 ? ns : 1) * dim);
  double* t = s;
  transform_points(&r, src, ns, s);
  if (!symmetric) {
    t = (double*)malloc(sizeof(double) * (nt > 0 ? nt : 1) * dim);
    transform_points(&r, trg, nt, t);
  } else {
    nt = ns;
  }
  double k0[MAXD * MAXD];
  double zero[MAXD] = {0, 0, 0};
  kernel_eval(&r, kind, zero, zero, k0);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < nt; ++i) {
    double acc[MAXD] = {0, 0, 0}, k[MAXD * MAXD];
    for (int64_t j = 0; j < ns; ++j) {
      if (symmetric && i == j) continue; /* full_direct.hpp:17-19 */
      kernel_eval(&r, kind, t + i * dim, s + j * dim, k);
      for (int b = 0; b < kn; ++b)
        for (int a = 0; a < km; ++a) acc[b] += w[j * km + a] * k[b * km + a];
    }
    if (symmetric) /* fmm_symmetric_evaluator.hpp:163-193 */
      for (int b = 0; b < kn; ++b)
        for (int a = 0; a < km; ++a) acc[b] += w[i * km + a] * k0[b * km + a];
    for (int b = 0; b < kn; ++b) out[i * kn + b] = acc[b];
  }
  if (t != s) free(t);
  free(s);
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * Interpolator tables (see header comment).
 * ------------------------------------------------------------------------------------------ */
static double node_pos(int i, int p) { return -1.0 + 2.0 * i / (p - 1); }

static void bary_weights(int p, int d, double* beta) {
  const int n = p - 1;
  if (d < 0 || d > n) d = n; /* kClassic: polynomial interpolant == FH of degree n */
  for (int k = 0; k <= n; ++k) {
    double s = 0.0;
    int lo = k - d > 0 ? k - d : 0, hi = k < n - d ? k : n - d;
    for (int i = lo; i <= hi; ++i) {
      double c = 1.0; /* C(d, k - i) */
      int b = k - i;
      for (int q = 1; q <= b; ++q) c = c * (d - b + q) / q;
      s += c;
    }
    beta[k] = ((k - d) % 2 == 0 ? 1.0 : -1.0) * s;
  }
}

static void bary_basis(int p, const double* beta, double t, double* s) {
  int hit = -1;
  double sum = 0.0;
  for (int i = 0; i < p; ++i) {
    double dt = t - node_pos(i, p);
    if (dt == 0.0) hit = i;
    s[i] = beta[i] / dt;
    sum += s[i];
  }
  if (hit >= 0) for (int i = 0; i < p; ++i) s[i] = i == hit ? 1.0 : 0.0;
  else for (int i = 0; i < p; ++i) s[i] /= sum;
}

static int ipow(int b, int e) { int r = 1; while (e-- > 0) r *= b; return r; }

/* ------------------------------------------------------------------------------------------
 * Small dense DFTs on the (2p-1)^dim circulant embedding, half spectrum on the last axis.
 * ------------------------------------------------------------------------------------------ */
typedef struct { double re, im; } cplx;

/* out[o][k][i] = sum_n in[o][n][i] W^(sgn k n), in complex */
static void dft_axis(const cplx* in, cplx* out, int outer, int n_in, int n_out, int inner, const cplx* tw, int nf,
                     int conj) {
  for (int o = 0; o < outer; ++o)
    for (int k = 0; k < n_out; ++k)
      for (int i = 0; i < inner; ++i) {
        double re = 0.0, im = 0.0;
        int idx = 0;
        for (int n = 0; n < n_in; ++n) {
          cplx v = in[((size_t)o * n_in + n) * inner + i];
          double wr = tw[idx].re, wi = conj ? -tw[idx].im : tw[idx].im;
          re += v.re * wr - v.im * wi;
          im += v.re * wi + v.im * wr;
          idx += k; if (idx >= nf) idx -= nf;
        }
        out[((size_t)o * n_out + k) * inner + i] = (cplx){re, im};
      }
}

/* Forward transform of a real array with n_in points per axis (zero padded to nf) to the half
 * spectrum [nf]^(dim-1) x [p].  n_in is p (multipoles) or nf (M2L operator). */
static void fwd_dft(const double* in, int dim, int n_in, int p, int nf, const cplx* tw, cplx* out, cplx* tmpA,
                    cplx* tmpB) {
  int outer = ipow(n_in, dim - 1);
  /* last axis: real -> p frequencies */
  cplx* first = dim == 1 ? out : tmpA;
  for (int o = 0; o < outer; ++o)
    for (int k = 0; k < p; ++k) {
      double re = 0.0, im = 0.0;
      int idx = 0;
      for (int n = 0; n < n_in; ++n) {
        double v = in[(size_t)o * n_in + n];
        re += v * tw[idx].re;
        im += v * tw[idx].im;
        idx += k; if (idx >= nf) idx -= nf;
      }
      first[(size_t)o * p + k] = (cplx){re, im};
    }
  cplx *src = tmpA, *dst = tmpB;
  int inner = p;
  outer = ipow(n_in, dim - 2 > 0 ? dim - 2 : 0);
  for (int a = dim - 2; a >= 0; --a) {
    cplx* o = a == 0 ? out : dst;
    dft_axis(src, o, outer, n_in, nf, inner, tw, nf, 0);
    cplx* t = src; src = dst; dst = t;
    inner *= nf;
    outer /= n_in;
  }
}

/* Inverse transform of a half spectrum, pruned to the first p points per axis. */
static void inv_dft(const cplx* in, int dim, int p, int nf, const cplx* tw, double* out, cplx* tmpA, cplx* tmpB) {
  const cplx* src = in;
  cplx* dst = tmpA;
  int outer = 1, inner = ipow(nf, dim - 2 > 0 ? dim - 2 : 0) * p;
  for (int a = 0; a + 1 < dim; ++a) {
    dft_axis(src, dst, outer, nf, p, inner, tw, nf, 1);
    src = dst;
    dst = dst == tmpA ? tmpB : tmpA;
    outer *= p;
    inner /= nf;
  }
  const int rows = ipow(p, dim - 1);
  for (int o = 0; o < rows; ++o)
    for (int m = 0; m < p; ++m) {
      double acc = src[(size_t)o * p].re;
      int idx = m;
      for (int k = 1; k < p; ++k) {
        cplx v = src[(size_t)o * p + k];
        /* Re(v e^{+i th}) with tw = (cos, -sin) */
        acc += 2.0 * (v.re * tw[idx].re + v.im * tw[idx].im);
        idx += m; if (idx >= nf) idx -= nf;
      }
      out[(size_t)o * p + m] = acc;
    }
}

/* ------------------------------------------------------------------------------------------
 * The FMM.  Uniform tree of `height` levels over the root box; dense row-major cell arrays.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int dim, height;
  int64_t n;
  double* pos;     /* sorted, transformed [n][dim] */
  int64_t* perm;   /* sorted -> caller index */
  int64_t* start;  /* leaf cell -> first point, [n_leaf_cells + 1] */
  unsigned char** occ; /* per level occupancy */
} orc_tree;

static int64_t cells_at(int dim, int level) { return (int64_t)1 << (dim * level); }

static int64_t cell_of(const double* x, int dim, int level, double width, const double* center) {
  const int nside = 1 << level;
  int64_t key = 0;
  for (int a = 0; a < dim; ++a) {
    int c = (int)floor((x[a] - (center[a] - 0.5 * width)) * nside / width);
    if (c < 0) c = 0;
    if (c >= nside) c = nside - 1;
    key = key * nside + c;
  }
  return key;
}

static void key_coords(int64_t key, int dim, int level, int* c) {
  const int nside = 1 << level;
  for (int a = dim - 1; a >= 0; --a) { c[a] = (int)(key % nside); key /= nside; }
}
static int64_t coords_key(const int* c, int dim, int level) {
  const int nside = 1 << level;
  int64_t key = 0;
  for (int a = 0; a < dim; ++a) key = key * nside + c[a];
  return key;
}

static void tree_build(orc_tree* t, int dim, int height, double width, const double* center, const double* pos,
                       int64_t n) {
  const int leaf = height - 1;
  const int64_t nc = cells_at(dim, leaf);
  t->dim = dim; t->height = height; t->n = n;
  int64_t* key = (int64_t*)malloc(sizeof(int64_t) * (n > 0 ? n : 1));
  t->start = (int64_t*)calloc(nc + 1, sizeof(int64_t));
  for (int64_t i = 0; i < n; ++i) {
    key[i] = cell_of(pos + i * dim, dim, leaf, width, center);
    t->start[key[i] + 1]++;
  }
  for (int64_t c = 0; c < nc; ++c) t->start[c + 1] += t->start[c];
  int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * nc);
  memcpy(fill, t->start, sizeof(int64_t) * nc);
  t->perm = (int64_t*)malloc(sizeof(int64_t) * (n > 0 ? n : 1));
  t->pos = (double*)malloc(sizeof(double) * (n > 0 ? n : 1) * dim);
  for (int64_t i = 0; i < n; ++i) { /* stable counting sort */
    int64_t s = fill[key[i]]++;
    t->perm[s] = i;
    memcpy(t->pos + s * dim, pos + i * dim, sizeof(double) * dim);
  }
  free(fill);
  free(key);
  t->occ = (unsigned char**)malloc(sizeof(unsigned char*) * height);
  for (int l = 0; l < height; ++l) t->occ[l] = (unsigned char*)calloc(cells_at(dim, l), 1);
  for (int64_t c = 0; c < nc; ++c)
    if (t->start[c + 1] > t->start[c]) {
      int cc[MAXD];
      key_coords(c, dim, leaf, cc);
      for (int l = leaf; l >= 0; --l) {
        t->occ[l][coords_key(cc, dim, l)] = 1;
        for (int a = 0; a < dim; ++a) cc[a] >>= 1;
      }
    }
}

static void tree_free(orc_tree* t) {
  for (int l = 0; l < t->height; ++l) free(t->occ[l]);
  free(t->occ); free(t->start); free(t->perm); free(t->pos);
}

static int tree_height_rule(int dim, int64_t n) { /* src/fmm/utility.hpp:12-16 */
  int h = (int)round(log((double)n) / log(pow(2.0, dim)));
  return h > 2 ? h : 2;
}

/* per-axis p x p contraction on an array viewed as [outer][p][inner] */
static void axis_contract(const double* in, double* out, int outer, int p, int inner, const double* tm, int transpose,
                          int accumulate) {
  for (int o = 0; o < outer; ++o)
    for (int r = 0; r < p; ++r)
      for (int i = 0; i < inner; ++i) {
        double acc = 0.0;
        for (int q = 0; q < p; ++q)
          acc += (transpose ? tm[q * p + r] : tm[r * p + q]) * in[((size_t)o * p + q) * inner + i];
        size_t e = ((size_t)o * p + r) * inner + i;
        if (accumulate) out[e] += acc; else out[e] = acc;
      }
}

/* child (side per axis from the child's coordinates) <-> parent interpolation */
static void child_transfer(const double* in, double* out_acc, int dim, int p, const int* side, const double* child_tm,
                           int transpose, double* tmp0, double* tmp1) {
  const int P = ipow(p, dim);
  const double* src = in;
  double* bufs[2] = {tmp0, tmp1};
  int outer = 1, inner = P / p;
  for (int a = 0; a < dim; ++a) {
    int last = a == dim - 1;
    double* dst = last ? out_acc : bufs[a & 1];
    axis_contract(src, dst, outer, p, inner, child_tm + side[a] * p * p, transpose, last);
    src = dst;
    outer *= p;
    inner /= p;
  }
}

/* Evaluate the FMM.  Points in ORIGINAL coordinates, bbox in original coordinates.
 * tree_height <= 0: src/fmm/utility.hpp rule on max(ns, nt) (generic) or ns (symmetric).
 * symmetric: targets = sources; the i == j pair is the k(0,0) w_i self term.
 * Returns 0, or -1 if the configuration is degenerate (height < 2). */
int orc_fmm(int rbf_id, int part, int dim, const double* params, const double* aniso, int kind,
            const double* bbox_min, const double* bbox_max, const double* src, int64_t ns, const double* trg,
            int64_t nt, const double* w, int symmetric, int order, int d, int tree_height, double* out) {
  orc_rbf r = make_rbf(rbf_id, part, dim, params, aniso);
  const int km = kind_km(kind, dim), kn = kind_kn(kind, dim);
  if (symmetric) { trg = src; nt = ns; }
  if (tree_height <= 0) tree_height = tree_height_rule(dim, symmetric ? ns : (ns > nt ? ns : nt));
  if (tree_height < 2) return -1;
  const int p = order, nf = 2 * p - 1, P = ipow(p, dim), F = ipow(nf, dim - 1) * p, leaf = tree_height - 1;
  const int NC = 1 << dim, NOFF = ipow(7, dim), NT = ipow(nf, dim);

  /* root box: src/fmm/utility.hpp:18-33 */
  double lo[MAXD], hi[MAXD], center[MAXD], width = 0.0;
  for (int a = 0; a < dim; ++a) { lo[a] = INFINITY; hi[a] = -INFINITY; }
  for (int c = 0; c < (1 << dim); ++c) {
    double pt[MAXD];
    for (int b = 0; b < dim; ++b) pt[b] = ((c >> b) & 1) ? bbox_max[b] : bbox_min[b];
    for (int a = 0; a < dim; ++a) {
      double s = 0.0;
      for (int b = 0; b < dim; ++b) s += pt[b] * r.A[a * dim + b];
      if (s < lo[a]) lo[a] = s;
      if (s > hi[a]) hi[a] = s;
    }
  }
  for (int a = 0; a < dim; ++a) { if (hi[a] - lo[a] > width) width = hi[a] - lo[a]; center[a] = lo[a] + 0.5 * (hi[a] - lo[a]); }
  width *= 1.01;
  if (width == 0.0) width = 1.0;

  /* trees */
  double* spos = (double*)malloc(sizeof(double) * (ns > 0 ? ns : 1) * dim);
  transform_points(&r, src, ns, spos);
  orc_tree st, tt_own, *tt = &st;
  tree_build(&st, dim, tree_height, width, center, spos, ns);
  free(spos);
  if (!symmetric) {
    double* tpos = (double*)malloc(sizeof(double) * (nt > 0 ? nt : 1) * dim);
    transform_points(&r, trg, nt, tpos);
    tree_build(&tt_own, dim, tree_height, width, center, tpos, nt);
    free(tpos);
    tt = &tt_own;
  }

  /* interpolator tables */
  double beta[64];
  bary_weights(p, d, beta);
  double* child_tm = (double*)malloc(sizeof(double) * 2 * p * p); /* [side][parent node m][child node n] */
  {
    double s[64];
    for (int side = 0; side < 2; ++side)
      for (int n = 0; n < p; ++n) {
        double y = 0.5 * node_pos(n, p) + (side == 0 ? -0.5 : 0.5);
        bary_basis(p, beta, y, s);
        for (int m = 0; m < p; ++m) child_tm[(side * p + m) * p + n] = s[m];
      }
  }
  cplx* tw = (cplx*)malloc(sizeof(cplx) * nf);
  for (int j = 0; j < nf; ++j) {
    long double a = 6.283185307179586476925286766559L * j / nf;
    tw[j] = (cplx){(double)cosl(a), (double)-sinl(a)};
  }

  /* sorted weights */
  double* ws = (double*)malloc(sizeof(double) * (ns > 0 ? ns : 1) * km);
  for (int64_t i = 0; i < ns; ++i)
    for (int a = 0; a < km; ++a) ws[i * km + a] = w[st.perm[i] * km + a];

  /* (ORC_TIMING=1 in the environment prints the wall time of the three phases: the CPU baseline's own profile) */
  const int timing = getenv("ORC_TIMING") != NULL;
  const double t_up0 = omp_get_wtime();
  /* ---- upward: P2M at the leaves, M2M to level 2 ---- */
  double*** M = (double***)calloc(tree_height, sizeof(double**));
  cplx*** Mh = (cplx***)calloc(tree_height, sizeof(cplx**));
  for (int l = 2; l < tree_height; ++l) {
    M[l] = (double**)calloc(cells_at(dim, l), sizeof(double*));
    Mh[l] = (cplx**)calloc(cells_at(dim, l), sizeof(cplx*));
  }
  if (tree_height > 2) {
    const int64_t nleaf = cells_at(dim, leaf);
    const double cw = width / (1 << leaf);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t c = 0; c < nleaf; ++c) {
      if (!st.occ[leaf][c]) continue;
      double* Mc = (double*)calloc((size_t)km * P, sizeof(double));
      int cc[MAXD];
      key_coords(c, dim, leaf, cc);
      double basis[MAXD][64];
      for (int64_t i = st.start[c]; i < st.start[c + 1]; ++i) {
        for (int a = 0; a < dim; ++a) {
          double ctr = center[a] - 0.5 * width + (cc[a] + 0.5) * cw;
          bary_basis(p, beta, (st.pos[i * dim + a] - ctr) / (0.5 * cw), basis[a]);
        }
        for (int n = 0; n < P; ++n) {
          int rr = n;
          double s = 1.0;
          for (int a = dim - 1; a >= 0; --a) { s *= basis[a][rr % p]; rr /= p; }
          for (int a = 0; a < km; ++a) Mc[a * P + n] += s * ws[i * km + a];
        }
      }
      M[leaf][c] = Mc;
    }
    for (int l = leaf - 1; l >= 2; --l) {
      const int64_t nc = cells_at(dim, l);
#pragma omp parallel for schedule(dynamic, 16)
      for (int64_t c = 0; c < nc; ++c) {
        if (!st.occ[l][c]) continue;
        double* Mc = (double*)calloc((size_t)km * P, sizeof(double));
        double* t0 = (double*)malloc(sizeof(double) * P * 2);
        int cc[MAXD];
        key_coords(c, dim, l, cc);
        for (int ch = 0; ch < NC; ++ch) {
          int c2[MAXD], side[MAXD];
          for (int a = 0; a < dim; ++a) { side[a] = (ch >> (dim - 1 - a)) & 1; c2[a] = 2 * cc[a] + side[a]; }
          const double* Mch = M[l + 1][coords_key(c2, dim, l + 1)];
          if (!Mch) continue;
          for (int a = 0; a < km; ++a) child_transfer(Mch + a * P, Mc + a * P, dim, p, side, child_tm, 0, t0, t0 + P);
        }
        free(t0);
        M[l][c] = Mc;
      }
    }
    /* multipole spectra */
    for (int l = 2; l < tree_height; ++l) {
      const int64_t nc = cells_at(dim, l);
#pragma omp parallel for schedule(dynamic, 16)
      for (int64_t c = 0; c < nc; ++c) {
        if (!M[l][c]) continue;
        cplx* h = (cplx*)malloc(sizeof(cplx) * km * F);
        cplx* tmp = (cplx*)malloc(sizeof(cplx) * 2 * (size_t)F);
        for (int a = 0; a < km; ++a) fwd_dft(M[l][c] + a * P, dim, p, p, nf, tw, h + (size_t)a * F, tmp, tmp + F);
        free(tmp);
        Mh[l][c] = h;
      }
    }
  }

  const double t_down0 = omp_get_wtime();
  /* ---- downward ---- */
  double*** L = (double***)calloc(tree_height, sizeof(double**));
  for (int l = 2; l < tree_height; ++l) L[l] = (double**)calloc(cells_at(dim, l), sizeof(double*));
  cplx* Kh = tree_height > 2 ? (cplx*)malloc(sizeof(cplx) * (size_t)NOFF * kn * km * F) : NULL;
  for (int l = 2; l < tree_height; ++l) {
    const double cw = width / (1 << l), h = cw / (p - 1);
    double scale = 1.0;
    for (int a = 0; a < dim; ++a) scale /= nf;
    /* M2L operators of this level (non-homogeneous kernels: per level) */
#pragma omp parallel for schedule(dynamic, 1)
    for (int oi = 0; oi < NOFF; ++oi) {
      int o[MAXD], rr = oi, near = 1;
      for (int a = dim - 1; a >= 0; --a) { o[a] = rr % 7 - 3; rr /= 7; if (o[a] < -1 || o[a] > 1) near = 0; }
      if (near) continue;
      double* T = (double*)malloc(sizeof(double) * (size_t)NT * kn * km);
      cplx* tmp = (cplx*)malloc(sizeof(cplx) * 2 * (size_t)F);
      for (int e = 0; e < NT; ++e) {
        double x[MAXD], y[MAXD] = {0, 0, 0}, k[MAXD * MAXD];
        int q = e;
        for (int a = dim - 1; a >= 0; --a) {
          int ia = q % nf; q /= nf;
          int delta = ia < p ? ia : ia - nf;
          x[a] = h * delta - cw * o[a]; /* x_m - y_n for source-minus-target cell offset o */
        }
        /* the kernel functors act on transformed positions; fold nothing: use the functor */
        kernel_eval(&r, kind, x, y, k);
        for (int c = 0; c < kn * km; ++c) T[(size_t)c * NT + e] = k[c] * scale;
      }
      for (int c = 0; c < kn * km; ++c)
        fwd_dft(T + (size_t)c * NT, dim, nf, p, nf, tw, Kh + ((size_t)oi * kn * km + c) * F, tmp, tmp + F);
      free(tmp);
      free(T);
    }
    const int64_t nc = cells_at(dim, l);
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t c = 0; c < nc; ++c) {
      if (!tt->occ[l][c]) continue;
      double* Lc = (double*)calloc((size_t)kn * P, sizeof(double));
      cplx* acc = (cplx*)calloc((size_t)kn * F, sizeof(cplx));
      cplx* tmp = (cplx*)malloc(sizeof(cplx) * 2 * (size_t)F);
      double* t0 = (double*)malloc(sizeof(double) * P * 2);
      int cc[MAXD], pc[MAXD];
      key_coords(c, dim, l, cc);
      for (int a = 0; a < dim; ++a) pc[a] = cc[a] >> 1;
      int any = 0;
      /* interaction list: children of the parent's neighbours that are not adjacent */
      const int nn = ipow(3, dim);
      for (int e = 0; e < nn; ++e) {
        int q[MAXD], rr = e, ok = 1;
        for (int a = dim - 1; a >= 0; --a) { q[a] = pc[a] + rr % 3 - 1; rr /= 3; if (q[a] < 0 || q[a] >= (1 << (l - 1))) ok = 0; }
        if (!ok) continue;
        for (int ch = 0; ch < NC; ++ch) {
          int s[MAXD], far = 0, oi = 0;
          for (int a = 0; a < dim; ++a) {
            s[a] = 2 * q[a] + ((ch >> (dim - 1 - a)) & 1);
            int o = s[a] - cc[a];
            if (o > 1 || o < -1) far = 1;
            oi = oi * 7 + o + 3;
          }
          if (!far) continue;
          const cplx* mh = Mh[l][coords_key(s, dim, l)];
          if (!mh) continue;
          any = 1;
          for (int b = 0; b < kn; ++b)
            for (int a = 0; a < km; ++a) {
              const cplx* kh = Kh + ((size_t)oi * kn * km + b * km + a) * F;
              const cplx* m = mh + (size_t)a * F;
              cplx* dst = acc + (size_t)b * F;
              for (int f = 0; f < F; ++f) {
                dst[f].re += kh[f].re * m[f].re - kh[f].im * m[f].im;
                dst[f].im += kh[f].re * m[f].im + kh[f].im * m[f].re;
              }
            }
        }
      }
      if (any)
        for (int b = 0; b < kn; ++b) inv_dft(acc + (size_t)b * F, dim, p, nf, tw, Lc + b * P, tmp, tmp + F);
      if (l > 2) { /* L2L from the parent */
        const double* Lp = L[l - 1][coords_key(pc, dim, l - 1)];
        int side[MAXD];
        for (int a = 0; a < dim; ++a) side[a] = cc[a] & 1;
        for (int b = 0; b < kn; ++b) child_transfer(Lp + b * P, Lc + b * P, dim, p, side, child_tm, 1, t0, t0 + P);
      }
      free(t0); free(tmp); free(acc);
      L[l][c] = Lc;
    }
  }
  free(Kh);

  const double t_leaf0 = omp_get_wtime();
  /* ---- leaves: L2P + P2P ---- */
  {
    const int64_t nleaf = cells_at(dim, leaf);
    const double cw = width / (1 << leaf);
    const int nn = ipow(3, dim);
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t c = 0; c < nleaf; ++c) {
      if (!tt->occ[leaf][c]) continue;
      int cc[MAXD];
      key_coords(c, dim, leaf, cc);
      const double* Lc = tree_height > 2 ? L[leaf][c] : NULL;
      double basis[MAXD][64];
      for (int64_t i = tt->start[c]; i < tt->start[c + 1]; ++i) {
        double v[MAXD] = {0, 0, 0};
        if (Lc) {
          for (int a = 0; a < dim; ++a) {
            double ctr = center[a] - 0.5 * width + (cc[a] + 0.5) * cw;
            bary_basis(p, beta, (tt->pos[i * dim + a] - ctr) / (0.5 * cw), basis[a]);
          }
          for (int n = 0; n < P; ++n) {
            int rr = n;
            double s = 1.0;
            for (int a = dim - 1; a >= 0; --a) { s *= basis[a][rr % p]; rr /= p; }
            for (int b = 0; b < kn; ++b) v[b] += s * Lc[b * P + n];
          }
        }
        for (int e = 0; e < nn; ++e) {
          int q[MAXD], rr = e, ok = 1;
          for (int a = dim - 1; a >= 0; --a) { q[a] = cc[a] + rr % 3 - 1; rr /= 3; if (q[a] < 0 || q[a] >= (1 << leaf)) ok = 0; }
          if (!ok) continue;
          int64_t sc = coords_key(q, dim, leaf);
          for (int64_t j = st.start[sc]; j < st.start[sc + 1]; ++j) {
            double k[MAXD * MAXD];
            /* symmetric: j == i is evaluated at d = 0, which is the k(0,0) w_i self term */
            kernel_eval(&r, kind, tt->pos + i * dim, st.pos + j * dim, k);
            for (int b = 0; b < kn; ++b)
              for (int a = 0; a < km; ++a) v[b] += ws[j * km + a] * k[b * km + a];
          }
        }
        for (int b = 0; b < kn; ++b) out[tt->perm[i] * kn + b] = v[b];
      }
    }
  }

  if (timing)
    fprintf(stderr, "orc_fmm: upward (P2M, M2M, forward DFTs) %.3f s, downward (operators, M2L, inverse DFTs, L2L) %.3f s, "
                    "leaves (L2P, P2P) %.3f s, %d threads\n",
            t_down0 - t_up0, t_leaf0 - t_down0, omp_get_wtime() - t_leaf0, omp_get_max_threads());
  for (int l = 2; l < tree_height; ++l) {
    for (int64_t c = 0; c < cells_at(dim, l); ++c) { free(M[l][c]); free(Mh[l][c]); free(L[l][c]); }
    free(M[l]); free(Mh[l]); free(L[l]);
  }
  free(M); free(Mh); free(L); free(ws); free(tw); free(child_tm);
  if (!symmetric) tree_free(&tt_own);
  tree_free(&st);
  return 0;
}

int orc_tree_height(int dim, int64_t n) { return tree_height_rule(dim, n); }

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
