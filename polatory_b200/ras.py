"""Restricted additive Schwarz preconditioner of the fit, on the device.

Restates, for value data and Hermite data (gradient points), one or several RBFs per model,
  preconditioner::RasPreconditioner   include/polatory/preconditioner/ras_preconditioner.hpp:34-364
  preconditioner::DomainDivider       include/polatory/preconditioner/domain_divider.hpp:17-321
  preconditioner::Domain              include/polatory/preconditioner/domain.hpp:16-52
  preconditioner::FineGrid            include/polatory/preconditioner/fine_grid.hpp:33-195
  preconditioner::CoarseGrid          include/polatory/preconditioner/coarse_grid.hpp:20-159
  preconditioner::mat_a               include/polatory/preconditioner/mat_a.hpp:10-61
  polynomial::UnisolventPointSet      include/polatory/polynomial/unisolvent_point_set.hpp:16-74
  polynomial::LagrangeBasis           include/polatory/polynomial/lagrange_basis.hpp:17-75

Host, once per fit: the level structure and the unisolvent points here (numpy); the choice of the coarse points
and the recursive bisection into overlapping domains in native multi-threaded code behind the C ABI
(csrc/ras_host.cu: `plt_ras_choose_coarse_points[_mixed]`, `plt_ras_divide_domains[_mixed]`).
Device: the Gram matrices of all domains of a level (batched kernels `plt_eval_gram_batched` /
`plt_eval_gram_mixed`), the reduction Q^T A Q, its batched Cholesky factorisation and the local solves of one
level as ONE launch of two triangular solves per domain -- hand-written kernels behind the C ABI
(csrc/ras_dense.cu: `plt_ras_reduce_q`, `plt_chol_batched`, `plt_chol_solve_batched`; no cuSOLVER / cuBLAS), with
the factors kept in HBM (the reference spills them to a temp file, preconditioner/binary_cache.hpp; 1M points
need ~19 GB here), and the level transfers
`update_residuals`, which are generic FMM evaluations (order 6, accuracy = infinity as the reference's
`Evaluator` default) whose trees / plans / operators stay resident between applications.
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import fmm
from .operator import monomial_basis

def _dense_lib():
    from . import _lib
    return _lib, _lib.load()


def reduce_q(a, q_top, out):
    """out[b] = Q^T A Q (fine_grid.hpp:71-81) on the device: a (B, m, m), q_top (B, l, m - l), out (B, m - l, m - l)."""
    import ctypes
    _lib, lib = _dense_lib()
    b, m = int(a.shape[0]), int(a.shape[1])
    l = int(q_top.shape[1])
    assert a.is_contiguous() and q_top.is_contiguous() and out.is_contiguous() and tuple(out.shape) == (b, m - l, m - l)
    st = lib.plt_ras_reduce_q(ctypes.c_void_p(a.data_ptr()), ctypes.c_void_p(q_top.data_ptr()), b, m, l,
                              ctypes.c_void_p(out.data_ptr()), None)
    if st != _lib.PLT_OK:
        raise RuntimeError("plt_ras_reduce_q failed")
    return out


def chol_batched(a, info):
    """In-place batched Cholesky (lower) of a (B, n, n); info (B,) int32: 0 or 1 + the first bad pivot."""
    import ctypes
    _lib, lib = _dense_lib()
    assert a.is_contiguous() and info.is_contiguous() and info.numel() == a.shape[0]
    st = lib.plt_chol_batched(ctypes.c_void_p(a.data_ptr()), int(a.shape[0]), int(a.shape[1]),
                              ctypes.c_void_p(info.data_ptr()), None)
    if st != _lib.PLT_OK:
        raise RuntimeError("plt_chol_batched failed")


def chol_solve_batched(factor, q_top, vals, lam):
    """lam[b] = [Q_top gamma; gamma] with L L^T gamma = vals[b, l:] + Q_top^T vals[b, :l] (fine_grid.hpp:112-133)."""
    import ctypes
    _lib, lib = _dense_lib()
    b, n = int(factor.shape[0]), int(factor.shape[1])
    l = 0 if q_top is None else int(q_top.shape[1])
    assert factor.is_contiguous() and vals.is_contiguous() and lam.is_contiguous()
    assert tuple(vals.shape) == (b, l + n) and tuple(lam.shape) == (b, l + n)
    st = lib.plt_chol_solve_batched(ctypes.c_void_p(factor.data_ptr()), b, n,
                                    ctypes.c_void_p(q_top.data_ptr()) if l else None, l,
                                    ctypes.c_void_p(vals.data_ptr()), ctypes.c_void_p(lam.data_ptr()), None)
    if st != _lib.PLT_OK:
        raise RuntimeError("plt_chol_solve_batched failed")
    return lam


def chol_inverse(factor):
    """(L L^T)^-1 of ONE factor (n, n) written by chol_batched: the solve kernel on the n unit vectors."""
    import ctypes
    import torch
    _lib, lib = _dense_lib()
    n = int(factor.shape[0])
    eye = torch.eye(n, dtype=torch.float64, device=factor.device)
    out = torch.empty_like(eye)
    st = lib.plt_chol_solve_shared(ctypes.c_void_p(factor.data_ptr()), n, n, ctypes.c_void_p(eye.data_ptr()),
                                   ctypes.c_void_p(out.data_ptr()), None)
    if st != _lib.PLT_OK:
        raise RuntimeError("plt_chol_solve_shared failed")
    return out


def gemv(a, x, out=None):
    """y = A x with the library's own kernel (A (rows, cols) row-major CUDA tensor)."""
    import ctypes
    import torch
    _lib, lib = _dense_lib()
    rows, cols = int(a.shape[0]), int(a.shape[1])
    x = x.contiguous()
    assert a.is_contiguous() and x.numel() == cols
    y = torch.empty(rows, dtype=torch.float64, device=a.device) if out is None else out
    st = lib.plt_gemv(ctypes.c_void_p(a.data_ptr()), rows, cols, ctypes.c_void_p(x.data_ptr()),
                      ctypes.c_void_p(y.data_ptr()), None)
    if st != _lib.PLT_OK:
        raise RuntimeError("plt_gemv failed")
    return y


K_FINE_TO_COARSE_RATIO = 10.0   # ras_preconditioner.hpp:53
K_N_COARSEST_POINTS = 2048      # ras_preconditioner.hpp:54
K_OVERLAP_QUOTA = 0.5           # domain_divider.hpp:25
K_MAX_LEAF_SIZE = 1024          # domain_divider.hpp:26


# ---------------------------------------------------------------------------------------------
# std::mt19937 (default seed) + libstdc++'s uniform_int_distribution (Lemire's method for 32-bit
# engines, bits/uniform_int_dist.h), as used by UnisolventPointSet.
# ---------------------------------------------------------------------------------------------
class _StdMt19937:
    def __init__(self, seed=5489):
        self._rs = np.random.RandomState(seed)  # init_genrand(seed): the std::mt19937 stream

    def __call__(self):
        return int.from_bytes(self._rs.bytes(4), "little")

    def uniform_index(self, n):
        """uniform_int_distribution<Index>(0, n - 1)(gen) for n <= 2^32."""
        product = self() * n
        low = product & 0xFFFFFFFF
        if low < n:
            threshold = ((1 << 32) - n) % n
            while low < threshold:
                product = self() * n
                low = product & 0xFFFFFFFF
        return product >> 32


def unisolvent_point_set(points, degree, dim):
    """UnisolventPointSet: the best-conditioned of 100 random candidate sets (sorted indices).
    Degree 0 is exact (every single point has rcond 1, the first trial wins); for higher degrees the
    reference ranks by Eigen's FullPivLU::rcond(), restated here as 1 / cond_1."""
    if degree < 0:
        return []
    n = len(points)
    from math import comb
    l = comb(dim + degree, degree)
    gen = _StdMt19937()
    best, best_rcond, found = None, 0.0, False
    for _ in range(100):
        s = set()
        while len(s) < l:
            s.add(gen.uniform_index(n))
        idx = sorted(s)
        p = monomial_basis(dim, degree, points[idx])
        try:
            rcond = 1.0 / np.linalg.cond(p, 1)
        except np.linalg.LinAlgError:
            continue
        if not np.isfinite(rcond) or rcond < 1e-15:
            continue
        found = True
        if best_rcond < rcond:
            best_rcond, best = rcond, idx
    if not found:
        raise RuntimeError("could not find a unisolvent set of points")
    return best


def lagrange_basis_matrix(points, poly_idcs, degree, dim):
    """LagrangeBasis(degree, points[poly_idcs]).evaluate(points): (mu x l)."""
    coeffs = np.linalg.inv(monomial_basis(dim, degree, points[poly_idcs]))
    return monomial_basis(dim, degree, points) @ coeffs


class Domain:
    __slots__ = ("point_indices", "inner_point")

    def __init__(self, point_indices, inner_point):
        self.point_indices = point_indices
        self.inner_point = inner_point


def divide_domains(a_points, point_idcs, poly_idcs, flat=False):
    """DomainDivider::divide_domains + Domain::merge_poly_points for value points only: the native
    multi-threaded implementation of the C ABI (csrc/ras_host.cu).  flat=True returns (offsets, indices,
    inner) instead of a list of Domain objects."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    a = np.ascontiguousarray(a_points, dtype=np.float64)
    idcs = np.ascontiguousarray(point_idcs, dtype=np.int64)
    poly = np.ascontiguousarray(poly_idcs, dtype=np.int64)
    h = ctypes.c_void_p()
    st = lib.plt_ras_divide_domains(a.ctypes.data, a.shape[1], idcs.ctypes.data, len(idcs), poly.ctypes.data, len(poly),
                                    K_MAX_LEAF_SIZE, K_OVERLAP_QUOTA, ctypes.byref(h))
    if st != _lib.PLT_OK:
        raise RuntimeError("plt_ras_divide_domains failed")
    try:
        n, total = lib.plt_ras_domains_count(h), lib.plt_ras_domains_total(h)
        off = np.empty(n + 1, dtype=np.int64)
        ind = np.empty(total, dtype=np.int64)
        inner = np.empty(total, dtype=np.uint8)
        lib.plt_ras_domains_get(h, off.ctypes.data, ind.ctypes.data, inner.ctypes.data)
    finally:
        lib.plt_ras_domains_destroy(h)
    if flat:
        return off, ind, inner.astype(bool)
    return [Domain(ind[off[i]:off[i + 1]], inner[off[i]:off[i + 1]].astype(bool)) for i in range(n)]


def choose_coarse_points(a_points, point_idcs, poly_idcs, n_coarse_points):
    """DomainDivider::choose_coarse_points (domain_divider.hpp:52-123): split the bounding-box
    clusters breadth-first (largest box first within a level) until there are n_coarse_points of
    them; the point nearest to each box centre is kept.  Native implementation (csrc/ras_host.cu)."""
    from . import _lib
    lib = _lib.load()
    a = np.ascontiguousarray(a_points, dtype=np.float64)
    idcs = np.ascontiguousarray(point_idcs, dtype=np.int64)
    poly = np.ascontiguousarray(poly_idcs, dtype=np.int64)
    out = np.empty(len(poly) + int(n_coarse_points), dtype=np.int64)
    st = lib.plt_ras_choose_coarse_points(a.ctypes.data, a.shape[1], idcs.ctypes.data, len(idcs), poly.ctypes.data,
                                          len(poly), int(n_coarse_points), out.ctypes.data)
    if st != _lib.PLT_OK:
        raise RuntimeError("plt_ras_choose_coarse_points failed")
    return out


def level_structure(n_rows):
    """Number of levels and the coarse point counts per level (ras_preconditioner.hpp:70-75,132-136)."""
    n_levels = max(int(math.ceil(math.log(n_rows / K_N_COARSEST_POINTS) / math.log(K_FINE_TO_COARSE_RATIO))), 0) + 1
    finest = math.log(n_rows) / math.log(K_FINE_TO_COARSE_RATIO)
    coarsest = math.log(K_N_COARSEST_POINTS) / math.log(K_FINE_TO_COARSE_RATIO)
    counts = {}
    for level in range(n_levels - 1, 0, -1):
        counts[level - 1] = int(K_FINE_TO_COARSE_RATIO ** (coarsest + (level - 1) * (finest - coarsest) / (n_levels - 1)))
    return n_levels, counts


# ---------------------------------------------------------------------------------------------
# Hermite data (gradient points, multiplicity `dim`): the same two algorithms with the reference's
# multiplicity-weighted cut ranks (domain_divider.hpp:205-231, 66-90), native as well (csrc/ras_host.cu).
# ---------------------------------------------------------------------------------------------
class MixedDomain:
    __slots__ = ("point_indices", "inner_point", "grad_point_indices", "inner_grad_point")


def divide_domains_mixed(a_points, a_grad_points, point_idcs, grad_idcs, poly_idcs):
    import ctypes
    from . import _lib
    lib = _lib.load()
    a = np.ascontiguousarray(a_points, dtype=np.float64)
    g = np.ascontiguousarray(a_grad_points, dtype=np.float64)
    pi = np.ascontiguousarray(point_idcs, dtype=np.int64)
    gi = np.ascontiguousarray(grad_idcs, dtype=np.int64)
    poly = np.ascontiguousarray(poly_idcs, dtype=np.int64)
    h = ctypes.c_void_p()
    st = lib.plt_ras_divide_domains_mixed(a.ctypes.data, g.ctypes.data, a.shape[1], pi.ctypes.data, len(pi),
                                          gi.ctypes.data, len(gi), poly.ctypes.data, len(poly), K_MAX_LEAF_SIZE,
                                          K_OVERLAP_QUOTA, ctypes.byref(h))
    if st != _lib.PLT_OK:
        raise RuntimeError("plt_ras_divide_domains_mixed failed")
    try:
        n, total, total_g = lib.plt_ras_domains_count(h), lib.plt_ras_domains_total(h), lib.plt_ras_domains_total_grads(h)
        off, ind, inner = np.empty(n + 1, dtype=np.int64), np.empty(total, dtype=np.int64), np.empty(total, dtype=np.uint8)
        offg, indg, innerg = np.empty(n + 1, dtype=np.int64), np.empty(total_g, dtype=np.int64), np.empty(total_g, dtype=np.uint8)
        lib.plt_ras_domains_get(h, off.ctypes.data, ind.ctypes.data, inner.ctypes.data)
        lib.plt_ras_domains_get_grads(h, offg.ctypes.data, indg.ctypes.data, innerg.ctypes.data)
    finally:
        lib.plt_ras_domains_destroy(h)
    out = []
    for i in range(n):
        d = MixedDomain()
        d.point_indices, d.inner_point = ind[off[i]:off[i + 1]], inner[off[i]:off[i + 1]].astype(bool)
        d.grad_point_indices, d.inner_grad_point = indg[offg[i]:offg[i + 1]], innerg[offg[i]:offg[i + 1]].astype(bool)
        out.append(d)
    return out


def choose_coarse_points_mixed(a_points, a_grad_points, point_idcs, grad_idcs, poly_idcs, n_coarse_rows):
    """choose_coarse_points with gradient points (a gradient centre counts `dim` rows)."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    a = np.ascontiguousarray(a_points, dtype=np.float64)
    g = np.ascontiguousarray(a_grad_points, dtype=np.float64)
    pi = np.ascontiguousarray(point_idcs, dtype=np.int64)
    gi = np.ascontiguousarray(grad_idcs, dtype=np.int64)
    poly = np.ascontiguousarray(poly_idcs, dtype=np.int64)
    outp = np.empty(len(poly) + int(n_coarse_rows) + 8, dtype=np.int64)
    outg = np.empty(int(n_coarse_rows) + 8, dtype=np.int64)
    np_, ng_ = ctypes.c_int64(), ctypes.c_int64()
    st = lib.plt_ras_choose_coarse_points_mixed(a.ctypes.data, g.ctypes.data, a.shape[1], pi.ctypes.data, len(pi),
                                                gi.ctypes.data, len(gi), poly.ctypes.data, len(poly), int(n_coarse_rows),
                                                outp.ctypes.data, ctypes.byref(np_), outg.ctypes.data, ctypes.byref(ng_))
    if st != _lib.PLT_OK:
        raise RuntimeError("plt_ras_choose_coarse_points_mixed failed")
    return outp[:np_.value].copy(), outg[:ng_.value].copy()


# ---------------------------------------------------------------------------------------------
# Device side
# ---------------------------------------------------------------------------------------------
class _FineLevel:
    """All FineGrids of one level, batched.  A domain is a list of ROWS of the global system (value row i =
    point i, rows mu + dim*j + c = component c of gradient point j), the l polynomial points first; padded to
    the level's largest domain; explicit inverses of Q^T A Q."""

    def __init__(self, ras, offsets, rows, inner):
        """offsets (n_dom + 1), rows (total) flat row indices domain after domain, inner (total) ownership flags."""
        torch = ras.torch
        dev, l = ras.device, ras.l
        self.n_dom = len(offsets) - 1
        cnt = np.diff(offsets).astype(np.int32)
        m_max = int(cnt.max())
        self.m = m_max
        r = m_max - l
        dom_of = np.repeat(np.arange(self.n_dom), cnt)
        col_of = np.arange(len(rows)) - np.repeat(offsets[:-1], cnt)
        idx = np.zeros((self.n_dom, m_max), dtype=np.int64)
        idx[dom_of, col_of] = rows
        inner = np.asarray(inner, dtype=bool)
        inner_glob = [np.asarray(rows)[inner]]
        inner_loc = [(dom_of * m_max + col_of)[inner]]
        self.idx = torch.from_numpy(idx).to(dev)
        self.cnt = torch.from_numpy(cnt).to(dev)
        self.valid = (torch.arange(m_max, device=dev)[None, :] < self.cnt[:, None])
        self.inner_glob = torch.from_numpy(np.concatenate(inner_glob)).to(dev)
        self.inner_loc = torch.from_numpy(np.concatenate(inner_loc)).to(dev)
        # Q = [q_top; I], q_top = -lagrange_p[rest]^T (fine_grid.hpp:71-73); padded rows are zero
        if l > 0:
            lag = ras.lagrange_p[self.idx]                      # (B, m, l)
            lag = lag * self.valid[:, :, None]
            self.q_top = -lag[:, l:, :].transpose(1, 2).contiguous()  # (B, l, r)
        else:
            self.q_top = None
        # Cholesky factors of Q^T A Q of every domain, kept in HBM (the reference spills its LDLT factors to a temp
        # file, binary_cache.hpp); hand-written batched kernels (csrc/ras_dense.cu), no host synchronisation
        self.fac = torch.empty((self.n_dom, r, r), dtype=torch.float64, device=dev)
        self.info = torch.zeros(self.n_dom, dtype=torch.int32, device=dev)
        chunk = max(1, min(self.n_dom, int(2 ** 31 // (8 * m_max * m_max))))
        for b0 in range(0, self.n_dom, chunk):
            b1 = min(self.n_dom, b0 + chunk)
            a = ras.gram(self.idx[b0:b1], self.cnt[b0:b1])      # (b, m, m)
            if l > 0:
                reduce_q(a, self.q_top[b0:b1], self.fac[b0:b1])
            else:
                self.fac[b0:b1].copy_(a)
            chol_batched(self.fac[b0:b1], self.info[b0:b1])
            del a

    def solve(self, ras, residuals, weights):
        """FineGrid::solve + set_solution_to for every domain of the level (fine_grid.hpp:103-147)."""
        vals = (residuals[self.idx] * self.valid).contiguous()   # (B, m), the l polynomial rows first
        lam = ras.torch.empty_like(vals)
        chol_solve_batched(self.fac, self.q_top, vals, lam)
        weights[self.inner_glob] = lam.reshape(-1)[self.inner_loc]


class _CoarseGrid:
    def __init__(self, ras, rows):
        torch = ras.torch
        dev, l = ras.device, ras.l
        self.idx = torch.from_numpy(np.asarray(rows, dtype=np.int64)).to(dev)
        m = len(rows)
        self.m = m
        cnt = torch.tensor([m], dtype=torch.int32, device=dev)
        a = ras.gram(self.idx[None], cnt)                          # (1, m, m)
        self.fac = torch.empty((1, m - l, m - l), dtype=torch.float64, device=dev)
        self.info = torch.zeros(1, dtype=torch.int32, device=dev)
        if l > 0:
            lag = ras.lagrange_p[self.idx]
            self.q_top = (-lag[l:, :].T).contiguous()[None]        # (1, l, m - l)
            reduce_q(a, self.q_top, self.fac)
            self.a_top = a[0, :l, :].clone()
            if ras.special_case:   # coarse_grid.hpp:75-77: the value point and the coarse grid's FIRST gradient point
                first_grad = (int(rows[1]) - ras.mu) // ras.dim
                p_top = monomial_basis(ras.dim, ras.model.poly_degree, ras.points[rows[:1]],
                                       ras.grad_points[first_grad:first_grad + 1])
            else:
                p_top = monomial_basis(ras.dim, ras.model.poly_degree, ras.points[rows[:l]])
            self.p_top_inv = torch.from_numpy(np.linalg.inv(p_top)).to(dev)   # l x l, host
        else:
            self.q_top = None
            self.fac.copy_(a)
        chol_batched(self.fac, self.info)
        # The coarse grid is solved ~2 n_levels times per application on ONE matrix: a single-CTA substitution is a
        # latency chain, so its inverse is formed once (the solve kernel on the unit vectors, all SMs busy) and
        # applied as a matrix-vector product.
        self.inv = chol_inverse(self.fac[0])

    def solve(self, ras, residuals, weights):
        """CoarseGrid::solve + set_solution_to (coarse_grid.hpp:84-128)."""
        torch = ras.torch
        l = ras.l
        vals = residuals[self.idx].contiguous()                    # (m,)
        if l > 0:
            q = self.q_top[0]                                      # (l, m - l)
            gamma = gemv(self.inv, vals[l:] + (q * vals[:l, None]).sum(dim=0))
            lam = torch.cat([(q * gamma[None, :]).sum(dim=1), gamma])
        else:
            lam = gemv(self.inv, vals)
        vals = vals[None]
        weights[self.idx] = lam
        if l > 0:
            # solve P c = d - A lambda at the polynomial points (l x l, l <= 10: element-wise products + sums)
            rhs = vals[0, :l] - (self.a_top * lam[None, :]).sum(dim=1)
            weights[ras.m_rows:] = (self.p_top_inv * rhs[None, :]).sum(dim=1)


class RasPreconditioner:
    """preconditioner::RasPreconditioner; `apply(v, out)` on CUDA tensors laid out as the operator's vectors
    [mu values | dim * sigma gradient components | l polynomial coefficients]."""

    def __init__(self, model, points, grad_points=None, device=None, verbose=False, transfer_config=None,
                 native_sweep=True):
        """transfer_config: optional (order, d) forced on the level-transfer evaluators, or "direct" for exact
        sums (default: the reference's accuracy = infinity, i.e. order 6) -- used by the parity tests to
        separate the FMM discretisation error of the transfers from everything else.
        native_sweep: apply() is ONE call into the library (plt_ras_sweep_apply, csrc/ras_sweep.cu); False keeps the
        torch re-expression of the same sweep below (the checker of the native one in the tests)."""
        import time
        import torch
        self.transfer_config = transfer_config
        self.torch = torch
        self.model = model
        dim = self.dim = model.dim
        self.l = model.poly_basis_size()
        self.points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, dim)
        self.grad_points = np.zeros((0, dim)) if grad_points is None else \
            np.ascontiguousarray(grad_points, dtype=np.float64).reshape(-1, dim)
        self.mu, self.sigma = len(self.points), len(self.grad_points)
        mu, sigma, l = self.mu, self.sigma, self.l
        self.m_rows = mu + dim * sigma
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        # "The special case" of ras_preconditioner.hpp:81-86 / coarse_grid.hpp:75-77: a linear polynomial pinned by
        # ONE value point and the first gradient point (test/interpolation/test_fitter.cpp:73).
        self.special_case = l > 0 and model.poly_degree == 1 and mu == 1 and sigma >= 1
        rbf = model.rbfs[0]
        self.bbox = fmm.Bbox.from_points(np.concatenate([self.points, self.grad_points]))
        # one handle per RBF: they carry the RBF constants for the Gram kernels (mat_a sums over the RBFs,
        # mat_a.hpp:23-55)
        self._gram_evs = [fmm.make_fmm_evaluator(r, self.bbox) for r in model.rbfs]
        n_levels, counts = level_structure(self.m_rows)
        self.n_levels = n_levels
        self.points_dev = torch.from_numpy(self.points).to(self.device)
        self.grad_points_dev = torch.from_numpy(self.grad_points).to(self.device)
        # per-ROW tables of the global system: coordinates of the row's point and the row type
        self.row_coords = torch.cat([self.points_dev, self.grad_points_dev.repeat_interleave(dim, dim=0)])
        self.row_types = torch.cat([torch.zeros(mu, dtype=torch.int8, device=self.device),
                                    (1 + torch.arange(dim, dtype=torch.int8, device=self.device)).repeat(sigma)])

        if self.special_case:
            poly_idcs = [0]
            coeffs = np.linalg.inv(monomial_basis(dim, model.poly_degree, self.points, self.grad_points[:1]))
        else:
            poly_idcs = unisolvent_point_set(self.points, model.poly_degree, dim) if l > 0 else []
            if l > 0:
                coeffs = np.linalg.inv(monomial_basis(dim, model.poly_degree, self.points[poly_idcs]))
        self.poly_idcs = poly_idcs
        if l > 0:
            self.lagrange_p = torch.from_numpy(
                monomial_basis(dim, model.poly_degree, self.points, self.grad_points) @ coeffs).to(self.device)
        point_idcs, grad_idcs = [None] * n_levels, [None] * n_levels
        rest = np.ones(mu, dtype=bool)
        rest[poly_idcs] = False
        point_idcs[n_levels - 1] = np.concatenate([np.asarray(poly_idcs, dtype=np.int64), np.nonzero(rest)[0]])
        grad_idcs[n_levels - 1] = np.arange(sigma, dtype=np.int64)
        aniso = np.asarray(rbf.anisotropy(), dtype=np.float64)
        # ras_preconditioner.hpp:111-120: the divider works on transformed points only for a single anisotropic RBF
        iso = len(model.rbfs) != 1 or np.array_equal(aniso, np.eye(dim))
        a_points = self.points if iso else self.points @ aniso.T
        a_grad_points = self.grad_points if iso else self.grad_points @ aniso.T
        self.fine = [None] * n_levels
        self.setup_seconds = {"coarse_points": 0.0, "divide_domains": 0.0, "factorize": 0.0}
        for level in range(n_levels - 1, 0, -1):
            # The domains of this level first: their factorisations are enqueued on the device and run while
            # the host chooses the coarse points of the next level (independent of the domains).
            t0 = time.perf_counter()
            if sigma == 0:
                d_off, d_rows, d_inner = divide_domains(a_points, point_idcs[level], poly_idcs, flat=True)
            else:
                doms = [(self._rows(d.point_indices, d.grad_point_indices),
                         np.concatenate([d.inner_point, np.repeat(d.inner_grad_point, dim)]))
                        for d in divide_domains_mixed(a_points, a_grad_points, point_idcs[level], grad_idcs[level],
                                                      poly_idcs)]
                d_off = np.concatenate([[0], np.cumsum([len(r_) for r_, _ in doms])]).astype(np.int64)
                d_rows = np.concatenate([r_ for r_, _ in doms])
                d_inner = np.concatenate([i_ for _, i_ in doms])
            n_domains = len(d_off) - 1
            t1 = time.perf_counter()
            self.fine[level] = _FineLevel(self, d_off, d_rows, d_inner)
            t2 = time.perf_counter()
            if sigma == 0:
                point_idcs[level - 1] = choose_coarse_points(a_points, point_idcs[level], poly_idcs, counts[level - 1])
                grad_idcs[level - 1] = np.zeros(0, dtype=np.int64)
            else:
                point_idcs[level - 1], grad_idcs[level - 1] = choose_coarse_points_mixed(
                    a_points, a_grad_points, point_idcs[level], grad_idcs[level], poly_idcs, counts[level - 1])
            t3 = time.perf_counter()
            self.setup_seconds["divide_domains"] += t1 - t0
            self.setup_seconds["factorize"] += t2 - t1   # host time to enqueue; the device part overlaps what follows
            self.setup_seconds["coarse_points"] += t3 - t2
            if verbose:
                print(f"level {level}: {n_domains} domains, {len(point_idcs[level])} points, {len(grad_idcs[level])} "
                      f"gradient points (domains {t1 - t0:.2f}s, factorisation enqueue {t2 - t1:.2f}s, coarse points "
                      f"{t3 - t2:.2f}s)", flush=True)
        self.point_idcs, self.grad_idcs = point_idcs, grad_idcs
        self.idx_dev = [torch.from_numpy(np.asarray(p, dtype=np.int64)).to(self.device) for p in point_idcs]
        self.gidx_dev = [torch.from_numpy(np.asarray(g, dtype=np.int64)).to(self.device) for g in grad_idcs]
        # flat rows of the gradient components of a level's gradient points, point-major (mu + dim*j + c)
        self.grows_dev = [(mu + dim * g[:, None] + torch.arange(dim, device=self.device)[None, :]).reshape(-1)
                          for g in self.gidx_dev]
        self.coarse = _CoarseGrid(self, self._rows(point_idcs[0], grad_idcs[0]))
        if verbose:
            print(f"level 0: 1 domain, {len(point_idcs[0])} points, {len(grad_idcs[0])} gradient points", flush=True)
        for f in self.fine + [self.coarse]:
            if f is not None and int(f.info.max()) != 0:
                raise RuntimeError("RAS: a local problem is not positive definite (duplicate points?)")
        self._evaluators = {}
        self.p = self.ap = None
        if l > 0:
            self.p_mono = torch.from_numpy(monomial_basis(dim, model.poly_degree, self.points, self.grad_points)).to(self.device)
        if n_levels > 1 and l > 0:
            # orthonormalised monomials and A p (ras_preconditioner.hpp:165-180)
            p = monomial_basis(dim, model.poly_degree, self.points, self.grad_points)
            for i in range(l):
                p[:, i] /= np.linalg.norm(p[:, i])
                for j in range(i + 1, l):
                    p[:, j] -= (p[:, i] @ p[:, j]) * p[:, i]
            self.p = torch.from_numpy(p).to(self.device)
            from .operator import Model as _Model, Operator
            # SymmetricEvaluator applies no nugget (ras_preconditioner.hpp:165-180): A p, not (A + nugget I) p
            fin = Operator(_Model(model.rbfs, poly_degree=-1, nugget=0.0), self.bbox, device=self.device)
            for ev in fin.a + fin.f + fin.ft + fin.h:
                self._configure_transfer(ev)
            fin.set_points(self.points, self.grad_points if sigma else None)
            self.ap = torch.empty_like(self.p)
            col = torch.empty(self.m_rows, dtype=torch.float64, device=self.device)
            for i in range(l):
                fin.apply(self.p[:, i].contiguous(), col)
                self.ap[:, i] = col
            del fin
        self._sweep = None
        if native_sweep and os.environ.get("PLT_RAS_PY_SWEEP") is None:
            self._build_native_sweep()

    def __del__(self):
        h = getattr(self, "_sweep", None)
        if h:
            self._sweep_lib.plt_ras_sweep_destroy(h)
            self._sweep = None

    def _build_native_sweep(self):
        """Hands the level structure to the library (device pointers into tensors this object keeps alive)."""
        import ctypes
        from . import _lib
        lib = _lib.load()
        ptr = lambda t: None if t is None or t.numel() == 0 else ctypes.c_void_p(t.data_ptr())  # noqa: E731
        n = self.n_levels
        h = ctypes.c_void_p()
        if lib.plt_ras_sweep_create(self.m_rows, self.l, n, ctypes.byref(h)) != _lib.PLT_OK:
            raise RuntimeError("plt_ras_sweep_create failed")
        self._sweep_lib = lib

        def ok(st):
            if st != _lib.PLT_OK:
                msg = lib.plt_ras_sweep_last_error(h)
                raise RuntimeError("RAS sweep: " + (msg.decode() if msg else f"status {st}"))

        keep = []
        for level in range(n):
            v, g = self.idx_dev[level].contiguous(), self.grows_dev[level].contiguous()
            keep += [v, g]
            ok(lib.plt_ras_sweep_set_level_rows(h, level, ptr(v), v.numel(), ptr(g), g.numel()))
        for level in range(1, n):
            f = self.fine[level]
            ok(lib.plt_ras_sweep_set_fine(h, level, f.n_dom, f.m, ptr(f.idx), ptr(f.cnt), ptr(f.fac), ptr(f.q_top),
                                          ptr(f.inner_glob), ptr(f.inner_loc), f.inner_glob.numel()))
        c = self.coarse
        if self.l > 0:
            q0 = c.q_top[0].contiguous()
            keep += [q0]
            ok(lib.plt_ras_sweep_set_coarse(h, c.m, ptr(c.idx), ptr(c.inv), ptr(q0), ptr(c.a_top), ptr(c.p_top_inv)))
        else:
            ok(lib.plt_ras_sweep_set_coarse(h, c.m, ptr(c.idx), ptr(c.inv), None, None, None))
        pairs = [(0, n - 1)] if n > 1 else []
        pairs += [(level, n - 1) for level in range(1, n - 1)]
        pairs += [(level, level - 1) for level in range(n - 1, 0, -1)]
        pairs += [(0, level - 1) for level in range(n - 1, 1, -1)]
        kind_of = {"a": 0, "f": 1, "ft": 2, "h": 3}
        for src, trg in dict.fromkeys(pairs):
            for name, ev, _fit in self._evaluator(src, trg):
                ok(lib.plt_ras_sweep_add_transfer(h, src, trg, kind_of[name], ev._h))
        if n > 1 and self.l > 0:
            ok(lib.plt_ras_sweep_set_poly(h, ptr(self.p_mono), ptr(self.p), ptr(self.ap)))
        self._sweep_keep = keep
        self._sweep = h

    # -- helpers ---------------------------------------------------------------------------
    def _configure_transfer(self, ev):
        if self.transfer_config == "direct":
            ev.force_direct(True)
        elif self.transfer_config:
            ev.force_config(*self.transfer_config)

    def _rows(self, point_indices, grad_point_indices):
        """Flat rows of the global system for a set of value and gradient points (values first)."""
        g = np.asarray(grad_point_indices, dtype=np.int64)
        grows = (self.mu + self.dim * g[:, None] + np.arange(self.dim)[None, :]).reshape(-1)
        return np.concatenate([np.asarray(point_indices, dtype=np.int64), grows])

    def gram(self, rows, counts):
        """mat_a for a batch of row sets: rows (B, m) flat row indices, counts (B,) valid rows -> (B, m, m);
        rows / columns beyond a set's count are identity."""
        torch = self.torch
        b, m = rows.shape
        out = torch.empty((b, m, m), dtype=torch.float64, device=self.device)
        pts = self.row_coords[rows].contiguous()
        if self.sigma > 0:
            types = self.row_types[rows]
            valid = torch.arange(m, device=self.device)[None, :] < counts[:, None]
            types = torch.where(valid, types, torch.full_like(types, -1)).contiguous()
        tmp = None
        for i, ev in enumerate(self._gram_evs):
            dst = out
            if i > 0:
                tmp = torch.empty_like(out) if tmp is None else tmp
                dst = tmp
            if self.sigma == 0:
                ev.gram_batched(pts, counts.contiguous(), self.model.nugget if i == 0 else 0.0, dst)
            else:
                ev.gram_mixed(pts, types, self.model.nugget if i == 0 else 0.0, dst)
            if i > 0:
                # rows / columns beyond a set's count are identity in every term: keep ONE identity
                pad = (torch.arange(m, device=self.device)[None, :] >= counts[:, None])
                tmp.diagonal(dim1=1, dim2=2)[pad] = 0.0
                out += tmp
        return out

    def _evaluator(self, src_level, trg_level):
        """interpolation::Evaluator(model, source level points) with the target level's points set
        (ras_preconditioner.hpp:251-265): the four kernel kinds, created for the non-empty blocks only."""
        key = (src_level, trg_level)
        if key not in self._evaluators:
            torch = self.torch
            sp = self.points_dev[self.idx_dev[src_level]].contiguous()
            sg = self.grad_points_dev[self.gidx_dev[src_level]].contiguous()
            tp = self.points_dev[self.idx_dev[trg_level]].contiguous()
            tg = self.grad_points_dev[self.gidx_dev[trg_level]].contiguous()
            evs = []
            for rbf in self.model.rbfs:       # evaluator.hpp:53-58: four kinds per RBF
                for name, make, s_pts, t_pts in (("a", fmm.make_fmm_evaluator, sp, tp),
                                                 ("f", fmm.make_fmm_gradient_evaluator, sg, tp),
                                                 ("ft", fmm.make_fmm_gradient_transpose_evaluator, sp, tg),
                                                 ("h", fmm.make_fmm_hessian_evaluator, sg, tg)):
                    if len(s_pts) == 0 or len(t_pts) == 0:
                        continue
                    ev = make(rbf, self.bbox)
                    self._configure_transfer(ev)
                    ev.set_source_points(s_pts)
                    ev.set_target_points(t_pts)
                    kn = self.dim if name in ("ft", "h") else 1
                    evs.append((name, ev, torch.empty(kn * len(t_pts), dtype=torch.float64, device=self.device)))
            self._evaluators[key] = evs
        return self._evaluators[key]

    def _solve(self, level, residuals):
        weights = self.torch.zeros(self.m_rows + self.l, dtype=self.torch.float64, device=self.device)
        if level == 0:
            self.coarse.solve(self, residuals, weights)
        else:
            self.fine[level].solve(self, residuals, weights)
        return weights

    def _update_residuals(self, src_level, trg_level, weights, residuals):
        """ras_preconditioner.hpp:287-321."""
        evs = self._evaluator(src_level, trg_level)
        w_v = weights[self.idx_dev[src_level]].contiguous()
        w_g = weights[self.grows_dev[src_level]].contiguous()
        trg_v, trg_g = self.idx_dev[trg_level], self.grows_dev[trg_level]
        src_w = {"a": w_v, "f": w_g, "ft": w_v, "h": w_g}
        trg_rows = {"a": trg_v, "f": trg_v, "ft": trg_g, "h": trg_g}
        for name, ev, fit in evs:
            ev.set_weights(src_w[name])
            ev.evaluate(fit)
            residuals[trg_rows[name]] -= fit
        if self.l > 0:
            c = weights[self.m_rows:]
            residuals[trg_v] -= self.p_mono[trg_v] @ c
            if len(trg_g):
                residuals[trg_g] -= self.p_mono[trg_g] @ c

    def _orthogonalize(self, weights, residuals):
        if self.l > 0:
            dot = self.p.T @ weights[:self.m_rows]
            weights[:self.m_rows] -= self.p @ dot
            residuals += self.ap @ dot

    # -- RasPreconditioner::operator() (ras_preconditioner.hpp:183-246) -----------------------
    def apply(self, v, out):
        if self._sweep:
            import ctypes
            assert v.is_contiguous() and out.is_contiguous() and v.numel() == self.m_rows + self.l == out.numel()
            st = self._sweep_lib.plt_ras_sweep_apply(self._sweep, ctypes.c_void_p(v.data_ptr()),
                                                     ctypes.c_void_p(out.data_ptr()), None)
            if st != 0:
                msg = self._sweep_lib.plt_ras_sweep_last_error(self._sweep)
                raise RuntimeError("RAS sweep: " + (msg.decode() if msg else f"status {st}"))
            return out
        return self.apply_reference(v, out)

    def apply_reference(self, v, out):
        """The same sweep as torch operations around the library's kernels (checker of the native sweep)."""
        n = self.n_levels
        residuals = v[:self.m_rows].clone()
        if n == 1:
            out.copy_(self._solve(0, residuals))
            return out
        total = self.torch.zeros_like(v)
        w = self._solve(0, residuals)
        self._update_residuals(0, n - 1, w, residuals)
        total += w
        for level in range(1, n - 1):
            w = self._solve(level, residuals)
            self._update_residuals(level, n - 1, w, residuals)
            total += w
            self._orthogonalize(total, residuals)
            w = self._solve(0, residuals)
            self._update_residuals(0, n - 1, w, residuals)
            total += w
        for level in range(n - 1, 0, -1):
            w = self._solve(level, residuals)
            self._update_residuals(level, level - 1, w, residuals)
            total += w
            self._orthogonalize(total, residuals)
            w = self._solve(0, residuals)
            if level > 1:
                self._update_residuals(0, level - 1, w, residuals)
            total += w
        out.copy_(total)
        return out

    def __call__(self, v):
        out = self.torch.empty_like(v)
        return self.apply(v, out)
