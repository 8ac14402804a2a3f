"""ORACLE (test infrastructure, never imported by the product path): numpy restatement of the reference's
RAS preconditioner for value data, dense and exact (every kernel sum is a direct sum).

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




def monomials(dim, degree, points):
    """polynomial::MonomialBasis::evaluate for value points (monomial_basis.hpp): 1 | x y z | x^2 xy xz y^2 yz z^2."""
    points = np.asarray(points, dtype=np.float64).reshape(-1, dim)
    cols = []
    if degree >= 0:
        cols.append(np.ones(len(points)))
    if degree >= 1:
        cols += [points[:, a] for a in range(dim)]
    if degree >= 2:
        cols += [points[:, a] * points[:, b] for a in range(dim) for b in range(a, dim)]
    return np.stack(cols, axis=1) if cols else np.zeros((len(points), 0))


class RasOracle:
    """Dense restatement of RasPreconditioner for one RBF and value data.  `a_dense[i, j]` = phi(x_i - x_j)
    over ALL points (anisotropy included, no nugget); `poly_idcs` are the unisolvent points to use."""

    def __init__(self, a_dense, points, dim, degree, nugget, poly_idcs):
        self.a_dense, self.points, self.dim, self.degree = a_dense, np.asarray(points, dtype=np.float64), dim, degree
        self.mu = len(self.points)
        self.l = monomials(dim, degree, self.points[:1]).shape[1]
        self.nugget = nugget
        mu, l = self.mu, self.l
        n_levels = max(int(math.ceil(math.log(mu / K_N_COARSEST_POINTS) / math.log(K_FINE_TO_COARSE_RATIO))), 0) + 1
        self.n_levels = n_levels
        poly_idcs = list(poly_idcs)
        if l:
            self.lagrange_p = monomials(dim, degree, self.points) @ np.linalg.inv(monomials(dim, degree, self.points[poly_idcs]))
        rest = np.ones(mu, dtype=bool)
        rest[poly_idcs] = False
        self.point_idcs = [None] * n_levels
        self.point_idcs[-1] = np.concatenate([np.asarray(poly_idcs, dtype=np.int64), np.nonzero(rest)[0]])
        finest = math.log(mu) / math.log(K_FINE_TO_COARSE_RATIO)
        coarsest = math.log(K_N_COARSEST_POINTS) / math.log(K_FINE_TO_COARSE_RATIO)
        self.fine = [None] * n_levels
        for level in range(n_levels - 1, 0, -1):
            n_coarse = int(K_FINE_TO_COARSE_RATIO ** (coarsest + (level - 1) * (finest - coarsest) / (n_levels - 1)))
            self.point_idcs[level - 1] = choose_coarse_points(self.points, self.point_idcs[level], poly_idcs, n_coarse)
            self.fine[level] = [self._grid(d.point_indices, d.inner_point)
                                for d in divide_domains(self.points, self.point_idcs[level], poly_idcs)]
        self.coarse = self._grid(self.point_idcs[0], None)
        if n_levels > 1 and l:
            p = monomials(dim, degree, self.points)
            for i in range(l):
                p[:, i] /= np.linalg.norm(p[:, i])
                for j in range(i + 1, l):
                    p[:, j] -= (p[:, i] @ p[:, j]) * p[:, i]
            self.p = p
            self.ap = self.a_dense @ p + nugget * p

    def _grid(self, idx, inner):
        l = self.l
        a = self.a_dense[np.ix_(idx, idx)] + self.nugget * np.eye(len(idx))  # mat_a
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
        w = np.zeros(self.mu + self.l)
        if level == 0:
            g = self.coarse
            lam = self._local(g, residuals)
            w[g["idx"]] = lam
            if self.l:
                l = self.l
                p_top = monomials(self.dim, self.degree, self.points[g["idx"][:l]])
                w[self.mu:] = np.linalg.solve(p_top, residuals[g["idx"][:l]] - g["a_top"] @ lam)
        else:
            for g in self.fine[level]:
                lam = self._local(g, residuals)
                w[g["idx"][g["inner"]]] = lam[g["inner"]]
        return w

    def _update(self, src, trg, w, residuals):
        si, ti = self.point_idcs[src], self.point_idcs[trg]
        fit = self.a_dense[np.ix_(ti, si)] @ w[si]
        if self.l:
            fit = fit + monomials(self.dim, self.degree, self.points[ti]) @ w[self.mu:]
        residuals[ti] -= fit

    def _orthogonalize(self, w, residuals):
        if self.l:
            dot = self.p.T @ w[:self.mu]
            w[:self.mu] -= self.p @ dot
            residuals += self.ap @ dot

    def __call__(self, v):
        n = self.n_levels
        residuals = np.array(v[:self.mu], dtype=np.float64)
        if n == 1:
            return self._solve(0, residuals)
        total = np.zeros(self.mu + self.l)
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
