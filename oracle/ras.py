"""ORACLE (test infrastructure, never imported by the product path): numpy restatement of the reference's
RAS preconditioner for value and Hermite data, dense and exact (every kernel sum is a direct sum).

  RasPreconditioner   include/polatory/preconditioner/ras_preconditioner.hpp:34-364
  DomainDivider       include/polatory/preconditioner/domain_divider.hpp:17-321
  Domain              include/polatory/preconditioner/domain.hpp:16-52
  FineGrid / CoarseGrid   include/polatory/preconditioner/fine_grid.hpp:33-195, coarse_grid.hpp:20-159
  mat_a               include/polatory/preconditioner/mat_a.hpp:10-61
  UnisolventPointSet / LagrangeBasis   include/polatory/polynomial/{unisolvent_point_set,lagrange_basis}.hpp

Parity unpinned against the reference itself (no golden vectors for the preconditioner in the reference's
tests, no reference build here); it pins the DEVICE implementation (batched explicit inverses, FMM level
transfers, native index bookkeeping) to the restated algorithm: same domains, same coarse points, FGMRES
iteration counts within +-1 (tests/test_gpu_ras.py, tests/test_ras_host.py).
"""
from __future__ import annotations

import heapq
import math

import numpy as np

K_FINE_TO_COARSE_RATIO = 10.0
K_N_COARSEST_POINTS = 2048
K_OVERLAP_QUOTA = 0.5
K_MAX_LEAF_SIZE = 1024


class Domain:
    def __init__(self, point_indices, inner_point):
        self.point_indices = point_indices
        self.inner_point = inner_point


def _round_half_to_even(d):
    return math.ceil((d - 0.5) / 2.0) + math.floor((d + 0.5) / 2.0)


def _sort_by_axes(pts):
    width = pts.max(axis=0) - pts.min(axis=0)
    axes = sorted(range(pts.shape[1]), key=lambda a: -width[a])
    return np.lexsort(tuple(pts[:, a] for a in reversed(axes)))


def divide_domains(a_points, point_idcs, poly_idcs):
    """DomainDivider::divide_domains + Domain::merge_poly_points (domain_divider.hpp:171-286, domain.hpp:33-51)
    for value points."""
    point_idcs = np.asarray(point_idcs, dtype=np.int64)
    queue = [Domain(point_idcs, np.ones(len(point_idcs), dtype=bool))]
    leaves = []
    head = 0
    while head < len(queue):
        d = queue[head]
        head += 1
        n = len(d.point_indices)
        if n <= K_MAX_LEAF_SIZE:
            leaves.append(d)
            continue
        order = _sort_by_axes(a_points[d.point_indices])
        idx, inner = d.point_indices[order], d.inner_point[order]
        q = K_OVERLAP_QUOTA * K_MAX_LEAF_SIZE / n
        n_sub = int(_round_half_to_even((1.0 + q) / 2.0 * n))
        left_part, right_part = n - n_sub, n_sub
        mid = int(_round_half_to_even((left_part + right_part) / 2.0))
        pos = np.arange(n)
        queue.append(Domain(idx[:right_part], inner[:right_part] & (pos[:right_part] < mid)))
        queue.append(Domain(idx[left_part:], inner[left_part:] & (pos[left_part:] >= mid)))
        queue[head - 1] = None
    poly = np.asarray(poly_idcs, dtype=np.int64)
    for d in leaves:  # merge_poly_points (domain.hpp:33-51)
        order = np.argsort(d.point_indices, kind="stable")
        idx, inner = d.point_indices[order], d.inner_point[order]
        if len(poly):
            pos = np.searchsorted(idx, poly)
            present = (pos < len(idx)) & (idx[np.minimum(pos, len(idx) - 1)] == poly)
            front_inner = np.zeros(len(poly), dtype=bool)
            front_inner[present] = inner[pos[present]]
            keep = np.ones(len(idx), dtype=bool)
            keep[pos[present]] = False
            idx = np.concatenate([poly, idx[keep]])
            inner = np.concatenate([front_inner, inner[keep]])
        d.point_indices, d.inner_point = idx, inner
    return leaves



def choose_coarse_points(a_points, point_idcs, poly_idcs, n_coarse_points):
    """DomainDivider::choose_coarse_points (domain_divider.hpp:52-123): the priority-queue walk as written."""
    poly_set = set(int(i) for i in poly_idcs)
    root = np.array([i for i in point_idcs if int(i) not in poly_set], dtype=np.int64)

    def init(idx):
        pts = a_points[idx]
        lo, hi = pts.min(axis=0), pts.max(axis=0)
        centre = 0.5 * (lo + hi)
        c = int(idx[np.argmin(((pts - centre) ** 2).sum(axis=1))])  # first minimum, as std::min_element
        return float(np.prod(hi - lo)), c, idx[_sort_by_axes(pts)]

    counter = 0
    vol, c, sorted_idx = init(root)
    heap = [(0, -vol, counter, c, sorted_idx)]
    while len(heap) < n_coarse_points:
        level, _, _, _, idx = heapq.heappop(heap)
        size = len(idx)
        if size % 2 == 0:
            mid = size // 2
        else:  # tie between (size-1)/2 and (size+1)/2: the even index wins (domain_divider.hpp:83-88)
            a = (size - 1) // 2
            mid = a if a % 2 == 0 else a + 1
            if size == 1:
                mid = 0
        for part in (idx[:mid], idx[mid:]):
            if len(part):
                counter += 1
                vol, c, s = init(part)
                heapq.heappush(heap, (level + 1, -vol, counter, c, s))
        if size == 1 and len(heap) >= len(root):
            break
    centres = []
    while heap:
        centres.append(heapq.heappop(heap)[3])
    return np.concatenate([np.asarray(poly_idcs, dtype=np.int64), np.asarray(centres, dtype=np.int64)])




def monomials(dim, degree, points, grad_points=None):
    """polynomial::MonomialBasis::evaluate (monomial_basis.hpp): value rows 1 | x y z | x^2 xy xz y^2 yz z^2, then
    `dim` derivative rows per gradient point (row mu + dim*j + k = d/dx_k at gradient point j)."""
    points = np.asarray(points, dtype=np.float64).reshape(-1, dim)
    exps = []
    if degree >= 0:
        exps.append((0,) * dim)
    if degree >= 1:
        exps += [tuple(1 if a == b else 0 for b in range(dim)) for a in range(dim)]
    if degree >= 2:
        for a in range(dim):
            for b in range(a, dim):
                e = [0] * dim
                e[a] += 1
                e[b] += 1
                exps.append(tuple(e))
    gp = np.zeros((0, dim)) if grad_points is None else np.asarray(grad_points, dtype=np.float64).reshape(-1, dim)
    mu, sigma = len(points), len(gp)
    out = np.zeros((mu + dim * sigma, len(exps)))
    for c, e in enumerate(exps):
        out[:mu, c] = np.prod(points ** np.asarray(e), axis=1)
        for k in range(dim):
            if e[k]:
                ek = list(e)
                ek[k] -= 1
                out[mu + k:mu + dim * sigma:dim, c] = e[k] * np.prod(gp ** np.asarray(ek), axis=1)
    return out


# -- the same two algorithms with gradient points (multiplicity dim), domain_divider.hpp:205-231, 66-90 ----------
def _mixed_coords(a_points, a_grad_points, idx, is_grad):
    out = np.empty((len(idx), a_points.shape[1]))
    out[~is_grad] = a_points[idx[~is_grad]]
    out[is_grad] = a_grad_points[idx[is_grad]]
    return out


def divide_domains_mixed(a_points, a_grad_points, point_idcs, grad_idcs, poly_idcs):
    """Returns a list of (point_indices, inner_point, grad_point_indices, inner_grad_point)."""
    dim = a_points.shape[1]
    idx0 = np.concatenate([np.asarray(point_idcs, dtype=np.int64), np.asarray(grad_idcs, dtype=np.int64)])
    g0 = np.concatenate([np.zeros(len(point_idcs), dtype=bool), np.ones(len(grad_idcs), dtype=bool)])
    queue, leaves, head = [(idx0, g0, np.ones(len(idx0), dtype=bool))], [], 0
    while head < len(queue):
        idx, isg, inner = queue[head]
        head += 1
        n_mult = int(np.where(isg, dim, 1).sum())
        if n_mult <= K_MAX_LEAF_SIZE:
            leaves.append((idx, isg, inner))
            continue
        order = _sort_by_axes(_mixed_coords(a_points, a_grad_points, idx, isg))
        idx, isg, inner = idx[order], isg[order], inner[order]
        prefix = np.concatenate([[0], np.cumsum(np.where(isg, dim, 1))])
        q = K_OVERLAP_QUOTA * K_MAX_LEAF_SIZE / n_mult
        n_sub = int(_round_half_to_even((1.0 + q) / 2.0 * n_mult))
        left_mult, right_mult = n_mult - n_sub, n_sub
        mid_mult = int(_round_half_to_even((left_mult + right_mult) / 2.0))
        ub = lambda x: int(np.searchsorted(prefix, x, side="right")) - 1
        left_part, right_part, mid = ub(left_mult), ub(right_mult), ub(mid_mult)
        pos = np.arange(len(idx))
        queue.append((idx[:right_part], isg[:right_part], inner[:right_part] & (pos[:right_part] < mid)))
        queue.append((idx[left_part:], isg[left_part:], inner[left_part:] & (pos[left_part:] >= mid)))
    out = []
    for idx, isg, inner in leaves:
        d = Domain(np.sort(idx[~isg]), None)
        order = np.argsort(idx[~isg], kind="stable")
        pi, pin = idx[~isg][order], inner[~isg][order]
        front = []
        for q_ in poly_idcs:  # merge_poly_points
            hit = np.nonzero(pi == q_)[0]
            front.append(bool(pin[hit[0]]) if len(hit) else False)
        keep = ~np.isin(pi, np.asarray(poly_idcs, dtype=np.int64))
        pi = np.concatenate([np.asarray(poly_idcs, dtype=np.int64), pi[keep]])
        pin = np.concatenate([np.asarray(front, dtype=bool), pin[keep]])
        out.append((pi, pin, idx[isg], inner[isg]))
    return out


def choose_coarse_points_mixed(a_points, a_grad_points, point_idcs, grad_idcs, poly_idcs, n_coarse_rows):
    dim = a_points.shape[1]
    poly_set = set(int(i) for i in poly_idcs)
    pv = np.array([i for i in point_idcs if int(i) not in poly_set], dtype=np.int64)
    idx0 = np.concatenate([pv, np.asarray(grad_idcs, dtype=np.int64)])
    g0 = np.concatenate([np.zeros(len(pv), dtype=bool), np.ones(len(grad_idcs), dtype=bool)])

    def init(idx, isg):
        pts = _mixed_coords(a_points, a_grad_points, idx, isg)
        lo, hi = pts.min(axis=0), pts.max(axis=0)
        k = int(np.argmin(((pts - 0.5 * (lo + hi)) ** 2).sum(axis=1)))
        order = _sort_by_axes(pts)
        return float(np.prod(hi - lo)), (int(idx[k]), bool(isg[k])), idx[order], isg[order]

    counter = 0
    vol, c, si, sg = init(idx0, g0)
    heap = [(0, -vol, counter, c, si, sg)]
    size = dim if c[1] else 1
    while size < n_coarse_rows:
        level, _, _, c, idx, isg = heapq.heappop(heap)
        size -= dim if c[1] else 1
        prefix = np.concatenate([[0], np.cumsum(np.where(isg, dim, 1))])
        d = np.abs(2 * prefix[:len(idx)] - int(prefix[-1]))
        cand = np.nonzero(d == d.min())[0]
        mid = int(cand[0])
        for k in cand[1:]:
            if k % 2 == 0:
                mid = int(k)
        for pi_, pg_ in ((idx[:mid], isg[:mid]), (idx[mid:], isg[mid:])):
            if len(pi_):
                counter += 1
                vol, c2, si, sg = init(pi_, pg_)
                size += dim if c2[1] else 1
                heapq.heappush(heap, (level + 1, -vol, counter, c2, si, sg))
    pts_out, grads_out = [int(i) for i in poly_idcs], []
    while heap:
        c = heapq.heappop(heap)[3]
        (grads_out if c[1] else pts_out).append(c[0])
    return np.asarray(pts_out, dtype=np.int64), np.asarray(grads_out, dtype=np.int64)


class RasOracle:
    """Dense restatement of RasPreconditioner for one RBF.  `a_dense` is the full matrix of the global system
    WITHOUT nugget and polynomial (mat_a over all rows: value row i = point i, rows mu + dim*j + c = component c
    of gradient point j; anisotropy included); `a_points` / `a_grad_points` are the coordinates the domain
    decomposition works on (anisotropy-transformed for one anisotropic RBF, ras_preconditioner.hpp:107-117);
    `poly_idcs` are the unisolvent points to use."""

    def __init__(self, a_dense, points, dim, degree, nugget, poly_idcs, grad_points=None, a_points=None,
                 a_grad_points=None):
        self.a_dense, self.dim, self.degree = a_dense, dim, degree
        self.points = np.asarray(points, dtype=np.float64)
        self.grad_points = np.zeros((0, dim)) if grad_points is None else np.asarray(grad_points, dtype=np.float64)
        a_points = self.points if a_points is None else a_points
        a_grad_points = self.grad_points if a_grad_points is None else a_grad_points
        self.mu, self.sigma = len(self.points), len(self.grad_points)
        mu, sigma = self.mu, self.sigma
        self.m_rows = mu + dim * sigma
        self.p_rows = monomials(dim, degree, self.points, self.grad_points)
        self.l = self.p_rows.shape[1]
        self.nugget = nugget
        l = self.l
        n_levels = max(int(math.ceil(math.log(self.m_rows / K_N_COARSEST_POINTS) / math.log(K_FINE_TO_COARSE_RATIO))), 0) + 1
        self.n_levels = n_levels
        poly_idcs = list(poly_idcs)
        self.poly_idcs = poly_idcs
        if l:
            self.lagrange_p = self.p_rows @ np.linalg.inv(monomials(dim, degree, self.points[poly_idcs]))
        rest = np.ones(mu, dtype=bool)
        rest[poly_idcs] = False
        self.point_idcs, self.grad_idcs = [None] * n_levels, [None] * n_levels
        self.point_idcs[-1] = np.concatenate([np.asarray(poly_idcs, dtype=np.int64), np.nonzero(rest)[0]])
        self.grad_idcs[-1] = np.arange(sigma, dtype=np.int64)
        finest = math.log(self.m_rows) / math.log(K_FINE_TO_COARSE_RATIO)
        coarsest = math.log(K_N_COARSEST_POINTS) / math.log(K_FINE_TO_COARSE_RATIO)
        self.fine = [None] * n_levels
        for level in range(n_levels - 1, 0, -1):
            n_coarse = int(K_FINE_TO_COARSE_RATIO ** (coarsest + (level - 1) * (finest - coarsest) / (n_levels - 1)))
            if sigma == 0:
                self.point_idcs[level - 1] = choose_coarse_points(a_points, self.point_idcs[level], poly_idcs, n_coarse)
                self.grad_idcs[level - 1] = np.zeros(0, dtype=np.int64)
                doms = [(d.point_indices, d.inner_point, np.zeros(0, dtype=np.int64), np.zeros(0, dtype=bool))
                        for d in divide_domains(a_points, self.point_idcs[level], poly_idcs)]
            else:
                self.point_idcs[level - 1], self.grad_idcs[level - 1] = choose_coarse_points_mixed(
                    a_points, a_grad_points, self.point_idcs[level], self.grad_idcs[level], poly_idcs, n_coarse)
                doms = divide_domains_mixed(a_points, a_grad_points, self.point_idcs[level], self.grad_idcs[level], poly_idcs)
            self.fine[level] = [self._grid(self._rows(pi, gi), np.concatenate([pin, np.repeat(gin, dim)]))
                                for pi, pin, gi, gin in doms]
        self.level_rows = [self._rows(self.point_idcs[k], self.grad_idcs[k]) for k in range(n_levels)]
        self.coarse = self._grid(self.level_rows[0], None)
        if n_levels > 1 and l:
            p = self.p_rows.copy()
            for i in range(l):
                p[:, i] /= np.linalg.norm(p[:, i])
                for j in range(i + 1, l):
                    p[:, j] -= (p[:, i] @ p[:, j]) * p[:, i]
            self.p = p
            # ap_ comes from SymmetricEvaluator, which applies no nugget (ras_preconditioner.hpp:165-180)
            self.ap = self.a_dense @ p

    def _rows(self, point_indices, grad_indices):
        g = np.asarray(grad_indices, dtype=np.int64)
        grows = (self.mu + self.dim * g[:, None] + np.arange(self.dim)[None, :]).reshape(-1)
        return np.concatenate([np.asarray(point_indices, dtype=np.int64), grows])

    def _grid(self, idx, inner):
        l = self.l
        a = self.a_dense[np.ix_(idx, idx)] + self.nugget * np.diag((np.asarray(idx) < self.mu).astype(float))  # mat_a
        g = {"idx": np.asarray(idx), "inner": inner, "a_top": a[:l]}
        if l:
            q_top = -self.lagrange_p[idx][l:].T
            g["q_top"] = q_top
            red = q_top.T @ a[:l, :l] @ q_top + q_top.T @ a[:l, l:] + a[l:, :l] @ q_top + a[l:, l:]
        else:
            red = a
        g["chol"] = np.linalg.cholesky(red)
        return g

    def _local(self, g, values):
        l = self.l
        d = values[g["idx"]]

        def chol_solve(rhs):
            y = np.linalg.solve(g["chol"], rhs)
            return np.linalg.solve(g["chol"].T, y)

        if l:
            gamma = chol_solve(g["q_top"].T @ d[:l] + d[l:])
            return np.concatenate([g["q_top"] @ gamma, gamma])
        return chol_solve(d)

    def _solve(self, level, residuals):
        w = np.zeros(self.m_rows + self.l)
        if level == 0:
            g = self.coarse
            lam = self._local(g, residuals)
            w[g["idx"]] = lam
            if self.l:
                l = self.l
                p_top = monomials(self.dim, self.degree, self.points[g["idx"][:l]])
                w[self.m_rows:] = np.linalg.solve(p_top, residuals[g["idx"][:l]] - g["a_top"] @ lam)
        else:
            for g in self.fine[level]:
                lam = self._local(g, residuals)
                w[g["idx"][g["inner"]]] = lam[g["inner"]]
        return w

    def _update(self, src, trg, w, residuals):
        si, ti = self.level_rows[src], self.level_rows[trg]
        fit = self.a_dense[np.ix_(ti, si)] @ w[si]
        if self.l:
            fit = fit + self.p_rows[ti] @ w[self.m_rows:]
        residuals[ti] -= fit

    def _orthogonalize(self, w, residuals):
        if self.l:
            dot = self.p.T @ w[:self.m_rows]
            w[:self.m_rows] -= self.p @ dot
            residuals += self.ap @ dot

    def __call__(self, v):
        n = self.n_levels
        residuals = np.array(v[:self.m_rows], dtype=np.float64)
        if n == 1:
            return self._solve(0, residuals)
        total = np.zeros(self.m_rows + self.l)
        w = self._solve(0, residuals)
        self._update(0, n - 1, w, residuals)
        total += w
        for level in range(1, n - 1):
            w = self._solve(level, residuals)
            self._update(level, n - 1, w, residuals)
            total += w
            self._orthogonalize(total, residuals)
            w = self._solve(0, residuals)
            self._update(0, n - 1, w, residuals)
            total += w
        for level in range(n - 1, 0, -1):
            w = self._solve(level, residuals)
            self._update(level, level - 1, w, residuals)
            total += w
            self._orthogonalize(total, residuals)
            w = self._solve(0, residuals)
            if level > 1:
                self._update(0, level - 1, w, residuals)
            total += w
        return total
