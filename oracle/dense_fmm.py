"""ORACLE (test infrastructure only) -- a second, independent restatement of the FMM discretisation.

Dense, FFT-free, pure numpy: every transfer is an explicit matrix built from its defining formula.

    P2M   M_c[n]   = sum_{j in c} S_n(x_j) w_j                   S_n = tensor product of 1-D barycentric bases
    M2M   M_p[m]  += sum_n S_m^p(y_n^c) M_c[n]                   y_n^c = node n of child c
    M2L   L_t[m]  += sum_n K(x_m^t - y_n^s) M_s[n]               s in the interaction list of t (levels >= 2)
    L2L   L_c[n]  += sum_m S_m^p(x_n^c) L_p[m]
    L2P   f(x_i)   = sum_m S_m(x_i) L_c[m]
    P2P   exact sums over the 3^dim adjacent leaves

on `order` equispaced nodes per axis (end points included), polynomial (d = -1) or Floater-Hormann (d >= 0)
barycentric weights, uniform 2^dim-tree of the given height, separation criterion 1 -- the discretisation
that oracle/fmm_oracle.c and the CUDA path implement with a circulant embedding + DFT for the M2L
(src/fmm/fmm_evaluator.hpp:57-58,85-99 for how the reference drives it; the engine itself is the un-vendored
polatory/ScalFMM3, see the header of fmm_oracle.c).  It shares no code with either: agreement of the FFT
paths with this file to ~1e-13 shows that the Fourier-space M2L (zero padding, twiddles, half spectrum,
pruned inverse) is an exact rewriting of the dense contraction, so that what stays "unpinned" against the
reference is only the published construction itself (node placement, FH weights, list definition).

Value kernel K only, isotropic or anisotropic; O(cells * P^2) memory-light loops: use small trees.
Only tests/ may import this.
"""
from __future__ import annotations

import itertools
from math import comb

import numpy as np

from . import rbf as orbf


def nodes(order):
    return -1.0 + 2.0 * np.arange(order) / (order - 1)


def bary_weights(order, d):
    n = order - 1
    if d < 0 or d > n:
        d = n
    beta = np.zeros(order)
    for k in range(order):
        s = sum(comb(d, k - i) for i in range(max(0, k - d), min(k, n - d) + 1))
        beta[k] = (-1.0) ** (k - d) * s
    return beta


def basis_1d(order, beta, t):
    """S_i(t) for an array of abscissae t: (len(t), order)."""
    t = np.atleast_1d(np.asarray(t, dtype=np.float64))
    x = nodes(order)
    diff = t[:, None] - x[None, :]
    hit = diff == 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        q = beta[None, :] / diff
        s = q / q.sum(axis=1, keepdims=True)
    rows = hit.any(axis=1)
    s[rows] = hit[rows].astype(np.float64)
    return s


def tensor_basis(order, beta, tpts):
    """Tensor-product basis at points tpts (n, dim) in the cell's [-1, 1]^dim coordinates: (n, order^dim),
    node index row-major (axis 0 slowest)."""
    n, dim = tpts.shape
    out = np.ones((n, 1))
    for a in range(dim):
        s = basis_1d(order, beta, tpts[:, a])
        out = (out[:, :, None] * s[:, None, :]).reshape(n, -1)
    return out


def fmm(name, params, dim, bbox_min, bbox_max, src, trg, w, order, d, height, aniso=None):
    rbf = orbf.make_rbf(name, params, dim, aniso)
    a = np.eye(dim) if aniso is None else np.asarray(aniso, dtype=np.float64)
    src_t = np.asarray(src, dtype=np.float64) @ a.T
    trg_t = np.asarray(trg, dtype=np.float64) @ a.T
    # src/fmm/utility.hpp:18-33: cube around the bbox of the transformed corners, width 1.01 x max extent
    corners = np.array([[bbox_max[b] if (c >> b) & 1 else bbox_min[b] for b in range(dim)] for c in range(1 << dim)])
    tc = corners @ a.T
    lo, hi = tc.min(axis=0), tc.max(axis=0)
    width = 1.01 * (hi - lo).max()
    center = lo + 0.5 * (hi - lo)
    origin = center - 0.5 * width
    beta = bary_weights(order, d)
    x1 = nodes(order)
    grid = np.array(list(itertools.product(x1, repeat=dim)))  # (P, dim), axis 0 slowest
    leaf = height - 1

    def phi(diff):
        return rbf.evaluate_isotropic(diff.reshape(-1, dim)).reshape(diff.shape[:-1])

    def cell_index(p, level):
        n = 1 << level
        c = np.floor((p - origin) / (width / n)).astype(np.int64)
        return np.clip(c, 0, n - 1)

    def cell_center(c, level):
        return origin + (np.asarray(c) + 0.5) * (width / (1 << level))

    def group(points, level):
        cells = {}
        ci = cell_index(points, level)
        for i, c in enumerate(map(tuple, ci)):
            cells.setdefault(c, []).append(i)
        return {c: np.asarray(v) for c, v in cells.items()}

    src_leaf = group(src_t, leaf)
    trg_leaf = group(trg_t, leaf)
    w = np.asarray(w, dtype=np.float64)

    # ---- upward
    M = [dict() for _ in range(height)]
    half = 0.5 * width / (1 << leaf)
    for c, idx in src_leaf.items():
        t = (src_t[idx] - cell_center(c, leaf)) / half
        M[leaf][c] = tensor_basis(order, beta, t).T @ w[idx]
    child_mats = {}
    for ch in itertools.product((0, 1), repeat=dim):
        # child nodes in the parent's coordinates
        y = 0.5 * grid + (np.asarray(ch) - 0.5)
        child_mats[ch] = tensor_basis(order, beta, y)  # (P_child_node, P_parent_node)
    for level in range(leaf - 1, 1, -1):
        for c, mc in M[level + 1].items():
            p = tuple(ci >> 1 for ci in c)
            ch = tuple(ci & 1 for ci in c)
            M[level][p] = M[level].get(p, 0.0) + child_mats[ch].T @ mc

    # ---- M2L (levels 2 .. leaf) + L2L
    L = [dict() for _ in range(height)]
    trg_cells = [set() for _ in range(height)]
    for c in trg_leaf:
        for level in range(leaf, 1, -1):
            trg_cells[level].add(tuple(ci >> (leaf - level) for ci in c))
    for level in range(2, height):
        n = 1 << level
        h = 0.5 * width / n
        for t in trg_cells[level]:
            acc = np.zeros(len(grid))
            xt = cell_center(t, level) + h * grid
            pt = [ci >> 1 for ci in t]
            for off in itertools.product(range(-1, 2), repeat=dim):
                pn = [pt[a_] + off[a_] for a_ in range(dim)]
                if any(q < 0 or q >= n // 2 for q in pn):
                    continue
                for ch in itertools.product((0, 1), repeat=dim):
                    s = tuple(2 * pn[a_] + ch[a_] for a_ in range(dim))
                    if max(abs(s[a_] - t[a_]) for a_ in range(dim)) <= 1:
                        continue  # adjacent: handled at a finer level / by P2P
                    ms = M[level].get(s)
                    if ms is None:
                        continue
                    ys = cell_center(s, level) + h * grid
                    acc += phi(xt[:, None, :] - ys[None, :, :]) @ ms
            if level > 2:
                p = tuple(pt)
                lp = L[level - 1].get(p)
                if lp is not None:
                    acc += child_mats[tuple(ci & 1 for ci in t)] @ lp
            L[level][t] = acc

    # ---- L2P + P2P
    out = np.zeros(len(trg_t))
    nleaf = 1 << leaf
    for c, idx in trg_leaf.items():
        if height > 2:
            t = (trg_t[idx] - cell_center(c, leaf)) / half
            out[idx] = tensor_basis(order, beta, t) @ L[leaf][c]
        for off in itertools.product(range(-1, 2), repeat=dim):
            s = tuple(c[a_] + off[a_] for a_ in range(dim))
            if any(q < 0 or q >= nleaf for q in s):
                continue
            sidx = src_leaf.get(s)
            if sidx is None:
                continue
            out[idx] += phi(trg_t[idx][:, None, :] - src_t[sidx][None, :, :]) @ w[sidx]
    return out
