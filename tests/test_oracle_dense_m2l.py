"""Second, independent check of the oracle's FMM discretisation ("tighten parity"): the C oracle's
Fourier-space M2L (circulant embedding, half spectrum, pruned inverse DFT), its M2M / L2L transfer matrices and
its interaction lists against oracle/dense_fmm.py, a pure-numpy FMM that applies every transfer as the dense
matrix of its defining formula  L_t[m] += sum_n K(x_m - y_n) M_s[n]  and shares no code with fmm_oracle.c.

Tolerance: 1e-13 * max|ref| at order 6-8 (measured 2e-15); 1e-11 at order 12, where the equispaced
interpolation amplifies FP64 rounding (two summation orders of the same formula differ by ~1e-12).
"""
import numpy as np
import pytest

from conftest import random_anisotropy
from oracle import dense_fmm
from oracle import fmm as ofmm


def _relerr(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


@pytest.mark.parametrize("dim,n,height,name,params,order,d,aniso,tol", [
    (3, 1200, 3, "bh3", [1.0, 0.0], 6, -1, False, 1e-13),    # one M2L level (the 2-level tree of the verdict)
    (3, 1500, 4, "exp", [1.1, 0.6], 6, -1, True, 1e-13),     # + M2M / L2L, anisotropic
    (3, 900, 3, "th3", [1.0, 0.05], 8, -1, False, 1e-13),
    (2, 1500, 4, "bh2", [1.0, 0.0], 12, 8, False, 1e-11),    # the matvec configuration (order 12, FH d = 8), 2-D
    (2, 1200, 5, "gc5", [1.0, 0.7], 10, -1, True, 1e-12),
    (1, 600, 5, "bh3", [1.0, 0.1], 12, 8, False, 1e-11),
])
def test_fft_m2l_equals_dense_contraction(dim, n, height, name, params, order, d, aniso, tol, rng):
    a = random_anisotropy(dim, rng) if aniso else None
    src = rng.uniform(-1, 1, (n, dim))
    trg = rng.uniform(-1, 1, (300, dim))
    w = rng.uniform(-1, 1, n)
    lo, hi = -np.ones(dim), np.ones(dim)
    dense = dense_fmm.fmm(name, params, dim, lo, hi, src, trg, w, order, d, height, a)
    fft = ofmm.fmm(name, params, dim, 0, lo, hi, src, trg, w, order, d, height, a)
    assert _relerr(fft, dense) < tol
    # and both are the FMM they claim to be: close to the exact sum, not identical to it
    exact = ofmm.direct(name, params, dim, 0, src, trg, w, a)
    assert 1e-13 < _relerr(dense, exact) < 1e-3


def test_fft_m2l_equals_dense_contraction_order12_3d(rng):
    """Order 12, d = 8 in 3-D (P = 1728) on a sparse tree: two source clusters, targets in two others, so that
    only a handful of dense 1728 x 1728 M2L blocks are formed."""
    dim, height = 3, 3
    src = np.concatenate([rng.uniform(-0.95, -0.55, (300, dim)), rng.uniform(0.55, 0.95, (300, dim))])
    trg = np.concatenate([rng.uniform(-0.95, -0.55, (60, dim)) * np.array([1, -1, 1]),
                          rng.uniform(0.05, 0.45, (60, dim))])
    w = rng.uniform(-1, 1, len(src))
    lo, hi = -np.ones(dim), np.ones(dim)
    dense = dense_fmm.fmm("bh3", [1.0, 0.0], dim, lo, hi, src, trg, w, 12, 8, height)
    fft = ofmm.fmm("bh3", [1.0, 0.0], dim, 0, lo, hi, src, trg, w, 12, 8, height)
    assert _relerr(fft, dense) < 1e-11
    exact = ofmm.direct("bh3", [1.0, 0.0], dim, 0, src, trg, w)
    assert _relerr(dense, exact) < 1e-8
