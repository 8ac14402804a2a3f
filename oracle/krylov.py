"""ORACLE (test infrastructure, never imported by the product path): numpy restatement of the
reference's GMRES / flexible GMRES.

  GmresBase  src/krylov/gmres_base.cpp:7-85 (setup, residuals, back substitution)
  Gmres      src/krylov/gmres.cpp:9-50      (classical Gram-Schmidt Arnoldi, Givens rotations)
  Fgmres     src/krylov/fgmres.cpp:8-28     (x = x0 + Z y)

Pinned by the reference's own acceptance test (test/krylov/test_krylov.cpp:84-110): after every
iteration the true relative residual equals the solver's estimate to 1e-12 and decreases
monotonically -- restated in tests/test_oracle_krylov.py.
"""
from __future__ import annotations

import numpy as np


class Fgmres:
    def __init__(self, op, rhs, max_iter):
        self.op = op
        self.rhs = np.asarray(rhs, dtype=np.float64)
        self.m = self.rhs.size
        self.max_iter = int(max_iter)
        self.x0 = np.zeros(self.m)
        self.right_pc = None
        self.iter = 0
        self.rhs_norm = float(np.linalg.norm(self.rhs))
        self.vs, self.zs = [], []

    def set_initial_solution(self, x0):
        self.x0 = np.array(x0, dtype=np.float64)

    def set_right_preconditioner(self, pc):
        self.right_pc = pc

    def setup(self):
        n = self.max_iter
        self.c = np.zeros(n)
        self.s = np.zeros(n)
        self.g = np.zeros(n + 1)
        r0 = self.rhs.copy() if not np.any(self.x0) else self.rhs - self.op(self.x0)
        self.g[0] = np.linalg.norm(r0)
        self.vs.append(r0 / self.g[0])
        self.r = np.zeros((n + 1, n))

    def iterate_process(self):
        if self.iter == self.max_iter:
            return
        j = self.iter
        z = self.right_pc(self.vs[j]) if self.right_pc is not None else self.vs[j].copy()
        self.zs.append(z)
        v = np.array(self.op(z), dtype=np.float64)
        for i in range(j + 1):
            self.r[i, j] = self.vs[i] @ v  # all against the un-updated vector (gmres.cpp:22-25)
        for i in range(j + 1):
            v = v - self.r[i, j] * self.vs[i]
        self.r[j + 1, j] = np.linalg.norm(v)
        self.vs.append(v / self.r[j + 1, j])
        r, c, s, g = self.r, self.c, self.s, self.g
        for i in range(j):
            x, y = r[i, j], r[i + 1, j]
            r[i, j] = c[i] * x + s[i] * y
            r[i + 1, j] = -s[i] * x + c[i] * y
        x, y = r[j, j], r[j + 1, j]
        den = np.hypot(x, y)
        c[j], s[j] = x / den, y / den
        r[j, j] = c[j] * x + s[j] * y
        g[j + 1] = -s[j] * g[j]
        g[j] = c[j] * g[j]
        self.iter += 1

    def solution_vector(self):
        it = self.iter
        y = np.zeros(it)
        for j in range(it - 1, -1, -1):
            y[j] = self.g[j]
            for i in range(j + 1, it):
                y[j] -= self.r[j, i] * y[i]
            y[j] /= self.r[j, j]
        x = self.x0.copy()
        for i in range(it):
            x += y[i] * self.zs[i]
        return x

    def iteration_count(self):
        return self.iter

    def max_iterations(self):
        return self.max_iter

    def absolute_residual(self):
        return abs(self.g[self.iter])

    def relative_residual(self):
        return abs(self.g[self.iter]) / self.rhs_norm
