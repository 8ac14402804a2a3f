"""ORACLE (test infrastructure only) -- exact O(N*M) direct summation in numpy.

Restates, for the four kernel kinds the FMM evaluator interface serves:
  * the kernel functors      include/polatory/fmm/kernel.hpp:43-52,
                             gradient_kernel.hpp:45-60, gradient_transpose_kernel.hpp:47-62,
                             hessian_kernel.hpp:46-65
  * the brute-force sums     src/fmm/full_direct.hpp:7-52
  * the self-interaction     src/fmm/fmm_symmetric_evaluator.hpp:163-193
  * the caller composition   include/polatory/interpolation/direct_evaluator.hpp:38-88

Nothing in the product path may import this module (see oracle/rbf.py header).
The reference holds no golden vectors for this path ("parity unpinned" against the
reference *FMM*; the direct sum is the exact quantity the reference's own tests compare
the FMM with, test/interpolation/test_evaluator.cpp:70-75).
"""
from __future__ import annotations

import numpy as np

KIND_K, KIND_F, KIND_FT, KIND_H = 0, 1, 2, 3


def kind_km(kind, dim):
    """Inputs per source point: kernel.hpp:29, gradient_kernel.hpp:28, ..."""
    return dim if kind in (KIND_F, KIND_H) else 1


def kind_kn(kind, dim):
    """Outputs per target point: kernel.hpp:30, gradient_transpose_kernel.hpp:30, ..."""
    return dim if kind in (KIND_FT, KIND_H) else 1


def transform_points(aniso, points):
    """geometry/point3d.hpp:43-47: points * A^T (row vectors)."""
    return np.asarray(points, dtype=np.float64) @ np.asarray(aniso, dtype=np.float64).T


def kernel_matrix(rbf, kind, diff_iso):
    """kernel.evaluate(x, y) on already-transformed positions; returns (n, kn, km).

    diff_iso = x - y in the isotropic (anisotropy-transformed) space.
    """
    a = rbf.aniso
    n = diff_iso.shape[0]
    if kind == KIND_K:  # kernel.hpp:43-52
        return rbf.evaluate_isotropic(diff_iso).reshape(n, 1, 1)
    if kind == KIND_F:  # gradient_kernel.hpp:45-60: -(grad_iso * A), km = Dim, kn = 1
        g = rbf.evaluate_gradient_isotropic(diff_iso) @ a
        return (-g)[:, None, :]
    if kind == KIND_FT:  # gradient_transpose_kernel.hpp:47-62: +(grad_iso * A), kn = Dim
        g = rbf.evaluate_gradient_isotropic(diff_iso) @ a
        return g[:, :, None]
    if kind == KIND_H:  # hessian_kernel.hpp:46-65: -(A^T H_iso A)
        h = rbf.evaluate_hessian_isotropic(diff_iso)
        return -np.einsum("ji,njk,kl->nil", a, h, a)
    raise ValueError("unknown kernel kind")


def full_direct(rbf, kind, src_points, trg_points, weights, symmetric=False, chunk=256):
    """src/fmm/full_direct.hpp:31-52 (generic) and :7-29 (symmetric, skips i == j and then
    adds k(0,0) w_i as fmm_symmetric_evaluator.hpp:163-193 does).

    Points are in ORIGINAL coordinates; the anisotropy is applied here the way
    set_source_points / set_target_points do (src/fmm/fmm_evaluator.hpp:120-157).
    weights: km per source, point-major.  Returns kn per target, point-major.
    """
    dim = rbf.dim
    km, kn = kind_km(kind, dim), kind_kn(kind, dim)
    src = transform_points(rbf.aniso, np.asarray(src_points, dtype=np.float64).reshape(-1, dim))
    trg = src if symmetric else transform_points(
        rbf.aniso, np.asarray(trg_points, dtype=np.float64).reshape(-1, dim))
    w = np.asarray(weights, dtype=np.float64).reshape(-1, km)
    assert w.shape[0] == src.shape[0]
    ns, nt = src.shape[0], trg.shape[0]
    out = np.zeros((nt, kn))
    if ns == 0 or nt == 0:
        return out.reshape(-1)
    for t0 in range(0, nt, chunk):
        t1 = min(nt, t0 + chunk)
        diff = (trg[t0:t1, None, :] - src[None, :, :]).reshape(-1, dim)
        with np.errstate(all="ignore"):
            k = kernel_matrix(rbf, kind, diff).reshape(t1 - t0, ns, kn, km)
        if symmetric:
            idx = np.arange(t0, t1)
            k[idx - t0, idx] = 0.0  # full_direct.hpp:17-19: skip src_idx == trg_idx
        out[t0:t1] = np.einsum("tsnm,sm->tn", k, w)
    if symmetric:
        with np.errstate(all="ignore"):
            k0 = kernel_matrix(rbf, kind, np.zeros((1, dim)))[0]  # evaluate(x, x)
        out += w @ k0.T
    return out.reshape(-1)


def direct_evaluator(rbf, nugget_unused, src_points, src_grad_points, weights,
                     trg_points, trg_grad_points):
    """include/polatory/interpolation/direct_evaluator.hpp:38-88 without the polynomial
    block (out of scope): y = [values at trg_points | Dim gradients at
    trg_grad_points], weights = [mu values | Dim*sigma gradient weights].
    """
    dim = rbf.dim
    mu = len(src_points)
    sigma = len(src_grad_points)
    w = np.asarray(weights, dtype=np.float64)
    wv, wg = w[:mu], w[mu:mu + dim * sigma]
    y_v = full_direct(rbf, KIND_K, src_points, trg_points, wv)
    y_v = y_v + full_direct(rbf, KIND_F, src_grad_points, trg_points, wg)
    y_g = full_direct(rbf, KIND_FT, src_points, trg_grad_points, wv)
    y_g = y_g + full_direct(rbf, KIND_H, src_grad_points, trg_grad_points, wg)
    return np.concatenate([y_v, y_g])
