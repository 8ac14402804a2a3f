"""Host-side RBF descriptors: the part of include/polatory/rbf/*.hpp (reference) the FMM
evaluator factories consume -- short name, parameters, anisotropy -- and nothing else.

The arithmetic of the RBFs lives in the CUDA kernels (polatory_b200/csrc/rbf.cuh); these
classes only carry the parameters across the C ABI, the way the reference's factories
`dynamic_cast` a runtime `Rbf<Dim>` to a concrete type (src/fmm/make_fmm_evaluator.cpp:40-69).
"""
from __future__ import annotations

import numpy as np

# include/polatory_b200.h PLT_RBF_*
_RBF_IDS = {
    "bh3": 0, "th3": 1, "bh2": 2, "th2": 3, "exp": 4, "gau": 5, "gc3": 6, "gc5": 7, "gc7": 8, "gc9": 9,
    "sp3": 10, "sp5": 11, "sp7": 12, "sp9": 13, "sph": 14, "cub": 15,
}
_CPD_ORDER = {"bh3": 1, "th3": 2, "bh2": 2, "th2": 3}
_POLYHARMONIC = ("bh3", "th3", "bh2", "th2")


class Rbf:
    """A runtime RBF (include/polatory/rbf/rbf.hpp, rbf_base.hpp:17-104)."""

    def __init__(self, short_name, params, dim, aniso=None):
        if short_name not in _RBF_IDS:
            # make_rbf.hpp:55
            raise RuntimeError(f"unknown RBF name: '{short_name}'")
        if dim not in (1, 2, 3):
            raise ValueError("dim must be 1, 2 or 3")
        self.short_name = short_name
        self.dim = dim
        params = [float(p) for p in params]
        if short_name in _POLYHARMONIC:
            # polyharmonic_odd.hpp:87-99
            if len(params) == 0:
                params = [1.0, 0.0]
            elif len(params) == 1:
                params = [params[0], 0.0]
        if len(params) != 2:
            # rbf_base.hpp:81-83
            raise ValueError("params.size() must be 2")
        self._params = params
        self._aniso = np.eye(dim)
        if aniso is not None:
            self.set_anisotropy(aniso)

    @property
    def rbf_id(self):
        return _RBF_IDS[self.short_name]

    def parameters(self):
        return list(self._params)

    def anisotropy(self):
        return self._aniso.copy()

    def set_anisotropy(self, aniso):
        aniso = np.asarray(aniso, dtype=np.float64).reshape(self.dim, self.dim)
        if not np.linalg.det(aniso) > 0.0:
            # rbf_base.hpp:73-75
            raise ValueError("aniso must have a positive determinant")
        self._aniso = aniso.copy()

    def cpd_order(self):
        return _CPD_ORDER.get(self.short_name, 0)

    def is_covariance_function(self):
        return self.short_name not in _POLYHARMONIC


def make_rbf(name, params, dim=3, aniso=None):
    """include/polatory/rbf/make_rbf.hpp:30-56."""
    return Rbf(name, params, dim, aniso)


def _named(short_name):
    def ctor(params=(), dim=3, aniso=None):
        return Rbf(short_name, params, dim, aniso)
    ctor.__name__ = short_name
    return ctor


Biharmonic3D = _named("bh3")
Triharmonic3D = _named("th3")
Biharmonic2D = _named("bh2")
Triharmonic2D = _named("th2")
CovExponential = _named("exp")
CovGaussian = _named("gau")
CovGeneralizedCauchy3 = _named("gc3")
CovGeneralizedCauchy5 = _named("gc5")
CovGeneralizedCauchy7 = _named("gc7")
CovGeneralizedCauchy9 = _named("gc9")
CovSpheroidal3 = _named("sp3")
CovSpheroidal5 = _named("sp5")
CovSpheroidal7 = _named("sp7")
CovSpheroidal9 = _named("sp9")
CovSpherical = _named("sph")
CovCubic = _named("cub")
