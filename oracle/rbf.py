"""ORACLE (test infrastructure only) -- numpy restatement of the reference RBF formulas.

Nothing in the product path (polatory_b200/) may import this module.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
and only as the checker.

Every function cites the reference file:line it follows (paths relative to
/root/reference).  All arithmetic is IEEE FP64, evaluated in the same operation order as
the reference where that is observable.

Parity status: the kernel *formulas* are pinned by the reference's own finite-difference
test (test/rbf/test_rbf.cpp:126-174, restated in tests/test_oracle_rbf.py).  The reference
holds no golden vectors for them.
"""
from __future__ import annotations

import numpy as np

# Short names in the order of make_rbf (include/polatory/rbf/make_rbf.hpp:30-45), plus the
# spheroidal direct / fast parts (include/polatory/rbf/cov_spheroidal3.hpp:113-123).
RBF_NAMES = [
    "bh3", "th3", "bh2", "th2", "exp", "gau", "gc3", "gc5", "gc7", "gc9",
    "sp3", "sp5", "sp7", "sp9", "sph", "cub",
]

FULL, DIRECT_PART, FAST_PART = 0, 1, 2


def _pow(x, n):
    # include/polatory/rbf/rbf_base.hpp:108-129 (pow<N>)
    if n == -1:
        return 1.0 / x
    if n == 0:
        return np.ones_like(x)
    if n == 1:
        return x
    if n == 2:
        return x * x
    if n == 3:
        return x * x * x
    if n == 4:
        x2 = x * x
        return x2 * x2
    return np.power(x, n)


def _sqrt_pow(x, n):
    # include/polatory/rbf/rbf_base.hpp:131-153 (sqrt_pow<N>)
    if n == 3:
        return x * np.sqrt(x)
    if n == 5:
        return x * x * np.sqrt(x)
    if n == 7:
        return x * x * x * np.sqrt(x)
    if n == 9:
        x2 = x * x
        return x2 * x2 * np.sqrt(x)
    if n == 11:
        x2 = x * x
        return x2 * x2 * x * np.sqrt(x)
    return np.power(x, n / 2.0)


def _outer(diff):
    # diff.transpose() * diff for row vectors: (n, D, D)
    return diff[:, :, None] * diff[:, None, :]


class Rbf:
    """Base: include/polatory/rbf/rbf_base.hpp:17-104."""

    short_name = ""
    cpd_order = 0
    part = FULL

    def __init__(self, params, dim, aniso=None):
        self.dim = dim
        self.params = self._default_params(list(params))
        self.aniso = np.eye(dim) if aniso is None else np.asarray(aniso, dtype=np.float64)
        if not np.linalg.det(self.aniso) > 0.0:
            # rbf_base.hpp:73-75
            raise ValueError("aniso must have a positive determinant")

    def _default_params(self, params):
        if len(params) != 2:
            raise ValueError("params.size() must be 2")
        return params

    def support_radius_isotropic(self):
        return np.inf

    # rbf_base.hpp:40-53 -- anisotropic wrappers; diff: (n, D) row vectors.
    def evaluate(self, diff):
        return self.evaluate_isotropic(diff @ self.aniso.T)

    def evaluate_gradient(self, diff):
        return self.evaluate_gradient_isotropic(diff @ self.aniso.T) @ self.aniso

    def evaluate_hessian(self, diff):
        h = self.evaluate_hessian_isotropic(diff @ self.aniso.T)
        return np.einsum("ji,njk,kl->nil", self.aniso, h, self.aniso)


class _PolyharmonicOdd(Rbf):
    """include/polatory/rbf/polyharmonic_odd.hpp:16-100."""

    K = 1

    def _default_params(self, params):
        # polyharmonic_odd.hpp:87-99
        if len(params) == 0:
            return [1.0, 0.0]
        if len(params) == 1:
            return [params[0], 0.0]
        if len(params) != 2:
            raise ValueError("params.size() must be 2")
        return params

    @property
    def sign(self):
        return 1.0 if ((self.K + 1) // 2) % 2 == 0 else -1.0  # :24

    @property
    def cpd_order(self):
        return (self.K + 1) // 2  # :33

    def _rho(self, diff):
        slope, c = self.params
        rho2 = np.sum(diff * diff, axis=1) + c * c
        return slope, rho2, np.sqrt(rho2)

    def evaluate_isotropic(self, diff):  # :32-39
        slope, rho2, rho = self._rho(diff)
        return self.sign * slope * _pow(rho, self.K)

    def evaluate_gradient_isotropic(self, diff):  # :41-53
        slope, rho2, rho = self._rho(diff)
        with np.errstate(divide="ignore", invalid="ignore"):
            coeff = self.sign * self.K * slope * _pow(rho, self.K - 2)
        coeff = np.where(rho == 0.0, 0.0, coeff)
        return coeff[:, None] * diff

    def evaluate_hessian_isotropic(self, diff):  # :55-67
        slope, rho2, rho = self._rho(diff)
        eye = np.eye(self.dim)[None]
        with np.errstate(divide="ignore", invalid="ignore"):
            coeff = self.sign * self.K * slope * _pow(rho, self.K - 2)
            h = coeff[:, None, None] * (eye + ((self.K - 2) / rho2)[:, None, None] * _outer(diff))
        return np.where((rho == 0.0)[:, None, None], 0.0, h)


class Biharmonic3D(_PolyharmonicOdd):
    short_name, K = "bh3", 1


class Triharmonic3D(_PolyharmonicOdd):
    short_name, K = "th3", 3


class _PolyharmonicEven(Rbf):
    """include/polatory/rbf/polyharmonic_even.hpp:16-107."""

    K = 2
    _default_params = _PolyharmonicOdd._default_params

    @property
    def sign(self):
        return 1.0 if (self.K // 2 + 1) % 2 == 0 else -1.0  # :24

    @property
    def cpd_order(self):
        return self.K // 2 + 1

    _rho = _PolyharmonicOdd._rho

    def evaluate_isotropic(self, diff):  # :33-44
        slope, rho2, rho = self._rho(diff)
        with np.errstate(divide="ignore", invalid="ignore"):
            v = self.sign * slope * _pow(rho, self.K) * np.log(rho)
        return np.where(rho == 0.0, 0.0, v)

    def evaluate_gradient_isotropic(self, diff):  # :46-58
        slope, rho2, rho = self._rho(diff)
        with np.errstate(divide="ignore", invalid="ignore"):
            coeff = self.sign * slope * _pow(rho, self.K - 2) * (1.0 + self.K * np.log(rho))
        coeff = np.where(rho == 0.0, 0.0, coeff)
        return coeff[:, None] * diff

    def evaluate_hessian_isotropic(self, diff):  # :60-73
        slope, rho2, rho = self._rho(diff)
        eye = np.eye(self.dim)[None]
        K = self.K
        with np.errstate(divide="ignore", invalid="ignore"):
            coeff = self.sign * slope * _pow(rho, K - 2) * (1.0 + K * np.log(rho))
            f = (K - 2.0 + K / (1.0 + K * np.log(rho))) / rho2
            h = coeff[:, None, None] * (eye + f[:, None, None] * _outer(diff))
        return np.where((rho == 0.0)[:, None, None], 0.0, h)


class Biharmonic2D(_PolyharmonicEven):
    short_name, K = "bh2", 2


class Triharmonic2D(_PolyharmonicEven):
    short_name, K = "th2", 4


class _Cov(Rbf):
    """include/polatory/rbf/covariance_function_base.hpp:10-39."""

    cpd_order = 0

    def _r(self, diff):
        psill, rng = self.params
        r = np.sqrt(np.sum(diff * diff, axis=1))  # diff.norm()
        return psill, rng, r, r / rng


class CovExponential(_Cov):
    """include/polatory/rbf/cov_exponential.hpp:33-61."""

    short_name = "exp"

    def evaluate_isotropic(self, diff):
        psill, rng, r, rho = self._r(diff)
        return psill * np.exp(-3.0 * rho)

    def evaluate_gradient_isotropic(self, diff):
        psill, rng, r, rho = self._r(diff)
        with np.errstate(divide="ignore", invalid="ignore"):
            coeff = -3.0 * psill * np.exp(-3.0 * rho) / (rng * r)
            return coeff[:, None] * diff

    def evaluate_hessian_isotropic(self, diff):
        psill, rng, r, rho = self._r(diff)
        eye = np.eye(self.dim)[None]
        with np.errstate(divide="ignore", invalid="ignore"):
            coeff = -3.0 * psill * np.exp(-3.0 * rho) / (rng * r)
            f = 1.0 / (r * r) + 3.0 / (rng * r)
            return coeff[:, None, None] * (eye - f[:, None, None] * _outer(diff))


class CovGaussian(_Cov):
    """include/polatory/rbf/cov_gaussian.hpp:33-60."""

    short_name = "gau"

    def evaluate_isotropic(self, diff):
        psill, rng, r, rho = self._r(diff)
        return psill * np.exp(-3.0 * rho * rho)

    def evaluate_gradient_isotropic(self, diff):
        psill, rng, r, rho = self._r(diff)
        coeff = -6.0 * psill * np.exp(-3.0 * rho * rho) / (rng * rng)
        return coeff[:, None] * diff

    def evaluate_hessian_isotropic(self, diff):
        psill, rng, r, rho = self._r(diff)
        eye = np.eye(self.dim)[None]
        coeff = -6.0 * psill * np.exp(-3.0 * rho * rho) / (rng * rng)
        return coeff[:, None, None] * (eye - 6.0 / (rng * rng) * _outer(diff))


class _CovGeneralizedCauchy(_Cov):
    """include/polatory/rbf/cov_generalized_cauchy3.hpp:24,35-63 (and 5, 7)."""

    kA = 7.0
    n = 3

    def evaluate_isotropic(self, diff):
        psill, rng, r, rho = self._r(diff)
        return psill / _sqrt_pow(1.0 + self.kA * rho * rho, self.n)

    def evaluate_gradient_isotropic(self, diff):
        psill, rng, r, rho = self._r(diff)
        coeff = -self.kA * float(self.n) * psill / (
            _sqrt_pow(1.0 + self.kA * rho * rho, self.n + 2) * rng * rng)
        return coeff[:, None] * diff

    def evaluate_hessian_isotropic(self, diff):
        psill, rng, r, rho = self._r(diff)
        eye = np.eye(self.dim)[None]
        coeff = -self.kA * float(self.n) * psill / (
            _sqrt_pow(1.0 + self.kA * rho * rho, self.n + 2) * rng * rng)
        f = self.kA * float(self.n + 2) / (self.kA * r * r + rng * rng)
        return coeff[:, None, None] * (eye - f[:, None, None] * _outer(diff))


class CovGeneralizedCauchy3(_CovGeneralizedCauchy):
    short_name, kA, n = "gc3", 7.0, 3


class CovGeneralizedCauchy5(_CovGeneralizedCauchy):
    short_name, kA, n = "gc5", 2.4822022531844965, 5


class CovGeneralizedCauchy7(_CovGeneralizedCauchy):
    short_name, kA, n = "gc7", 1.438027308408951, 7


class CovGeneralizedCauchy9(_Cov):
    """include/polatory/rbf/cov_generalized_cauchy9.hpp:33-59 (no kA factor)."""

    short_name = "gc9"

    def evaluate_isotropic(self, diff):
        psill, rng, r, rho = self._r(diff)
        return psill / _sqrt_pow(1.0 + rho * rho, 9)

    def evaluate_gradient_isotropic(self, diff):
        psill, rng, r, rho = self._r(diff)
        coeff = -9.0 * psill / (_sqrt_pow(1.0 + rho * rho, 11) * rng * rng)
        return coeff[:, None] * diff

    def evaluate_hessian_isotropic(self, diff):
        psill, rng, r, rho = self._r(diff)
        eye = np.eye(self.dim)[None]
        coeff = -9.0 * psill / (_sqrt_pow(1.0 + rho * rho, 11) * rng * rng)
        f = 11.0 / (r * r + rng * rng)
        return coeff[:, None, None] * (eye - f[:, None, None] * _outer(diff))


class _CovSpheroidal(_Cov):
    """include/polatory/rbf/cov_spheroidal3.hpp:27-128 (and 5, 7, 9)."""

    n = 3
    kRho0 = kA = kB = kC = kD = kE = 0.0

    def __init__(self, params, dim, aniso=None, part=FULL):
        super().__init__(params, dim, aniso)
        self.part = part

    def support_radius_isotropic(self):  # :108-111
        return self.kRho0 * self.params[1] if self.part == DIRECT_PART else np.inf

    def direct_part(self):  # :113-117
        return type(self)(self.params, self.dim, self.aniso, DIRECT_PART)

    def fast_part(self):  # :119-123
        return type(self)(self.params, self.dim, self.aniso, FAST_PART)

    def _select(self, rho, lin, imq, shape):
        m = (rho < self.kRho0).reshape((-1,) + (1,) * (len(shape) - 1))
        if self.part == DIRECT_PART:
            return np.where(m, lin - imq, 0.0)
        if self.part == FAST_PART:
            return imq
        return np.where(m, lin, imq)

    def evaluate_isotropic(self, diff):  # :44-60
        psill, rng, r, rho = self._r(diff)
        lin = psill * (1.0 - self.kA * rho)
        imq = psill * self.kB / _sqrt_pow(1.0 + self.kC * rho * rho, self.n)
        return self._select(rho, lin, imq, lin.shape)

    def evaluate_gradient_isotropic(self, diff):  # :62-84
        psill, rng, r, rho = self._r(diff)
        with np.errstate(divide="ignore", invalid="ignore"):
            lin = (-psill * self.kA / (r * rng))[:, None] * diff
        imq = (-psill * self.kD / (
            _sqrt_pow(1.0 + self.kC * rho * rho, self.n + 2) * rng * rng))[:, None] * diff
        return self._select(rho, lin, imq, lin.shape)

    def evaluate_hessian_isotropic(self, diff):  # :86-108
        psill, rng, r, rho = self._r(diff)
        eye = np.eye(self.dim)[None]
        o = _outer(diff)
        with np.errstate(divide="ignore", invalid="ignore"):
            cl = -psill * self.kA / (r * rng)
            lin = cl[:, None, None] * (eye - (1.0 / (r * r))[:, None, None] * o)
        ci = -psill * self.kD / (_sqrt_pow(1.0 + self.kC * rho * rho, self.n + 2) * rng * rng)
        f = float(self.n + 2) / (r * r + self.kE * rng * rng)
        imq = ci[:, None, None] * (eye - f[:, None, None] * o)
        return self._select(rho, lin, imq, lin.shape)


class CovSpheroidal3(_CovSpheroidal):
    short_name, n = "sp3", 3
    kRho0 = 0.18657871684006438
    kA = 2.009875543958482
    kB = 0.8734640537108553
    kC = 7.181510581693163
    kD = 18.81837403335934
    kE = 0.1392464703107397


class CovSpheroidal5(_CovSpheroidal):
    short_name, n = "sp5", 5
    kRho0 = 0.2580127411803573
    kA = 1.6149073288415876
    kB = 0.8575980168032007
    kC = 2.5036086535164204
    kD = 10.735449080535068
    kE = 0.39942344766841226


class CovSpheroidal7(_CovSpheroidal):
    short_name, n = "sp7", 7
    kRho0 = 0.2944149476843637
    kA = 1.4859979204216045
    kB = 0.8494862533016855
    kC = 1.44208314742683
    kD = 8.57520866899984
    kE = 0.6934412913598931


class CovSpheroidal9(_CovSpheroidal):
    # include/polatory/rbf/cov_spheroidal9.hpp:27-30,49,71,95-96: no kC / kE factors.
    short_name, n = "sp9", 9
    kRho0 = 0.31622776601683794
    kA = 1.4230249470757708
    kB = 0.8445585690332554
    kC = 1.0
    kD = 7.601027121299299
    kE = 1.0


class CovSpherical(_Cov):
    """include/polatory/rbf/cov_spherical.hpp:34-59."""

    short_name = "sph"

    def support_radius_isotropic(self):
        return self.params[1]

    def evaluate_isotropic(self, diff):
        psill, rng, r, rho = self._r(diff)
        return np.where(r < rng, psill * (1.0 + rho * (-1.5 + 0.5 * rho * rho)), 0.0)

    def evaluate_gradient_isotropic(self, diff):
        psill, rng, r, rho = self._r(diff)
        with np.errstate(divide="ignore", invalid="ignore"):
            coeff = np.where(r < rng, psill * (-1.5 / rho + 1.5 * rho) / (rng * rng), 0.0)
            return coeff[:, None] * diff

    def evaluate_hessian_isotropic(self, diff):
        raise RuntimeError("cov_spherical::evaluate_hessian_isotropic is not implemented")


class CovCubic(_Cov):
    """include/polatory/rbf/cov_cubic.hpp:34-64."""

    short_name = "cub"

    def support_radius_isotropic(self):
        return self.params[1]

    def evaluate_isotropic(self, diff):
        psill, rng, r, rho = self._r(diff)
        rho2 = rho * rho
        v = psill * (1.0 + rho2 * (-7.0 + rho * (8.75 + rho2 * (-3.5 + 0.75 * rho2))))
        return np.where(r < rng, v, 0.0)

    def evaluate_gradient_isotropic(self, diff):
        psill, rng, r, rho = self._r(diff)
        rho2 = rho * rho
        coeff = psill * (-14.0 + rho * (26.25 + rho2 * (-17.5 + 5.25 * rho2))) / (rng * rng)
        coeff = np.where(r < rng, coeff, 0.0)
        return coeff[:, None] * diff

    def evaluate_hessian_isotropic(self, diff):
        raise RuntimeError("cov_cubic::evaluate_hessian_isotropic is not implemented")


_CLASSES = {c.short_name: c for c in [
    Biharmonic3D, Triharmonic3D, Biharmonic2D, Triharmonic2D, CovExponential, CovGaussian,
    CovGeneralizedCauchy3, CovGeneralizedCauchy5, CovGeneralizedCauchy7, CovGeneralizedCauchy9,
    CovSpheroidal3, CovSpheroidal5, CovSpheroidal7, CovSpheroidal9, CovSpherical, CovCubic]}


def make_rbf(name, params, dim, aniso=None, part=FULL):
    """include/polatory/rbf/make_rbf.hpp:30-56."""
    if name not in _CLASSES:
        raise RuntimeError(f"unknown RBF name: '{name}'")
    cls = _CLASSES[name]
    if issubclass(cls, _CovSpheroidal):
        return cls(params, dim, aniso, part)
    if part != FULL:
        raise RuntimeError(f"'{name}' has no direct/fast split")
    return cls(params, dim, aniso)
