"""Host-side mirror of the reference's RBF operator (the Krylov matvec) over the device evaluators.

  interpolation::Operator<Dim>   include/polatory/interpolation/operator.hpp:18-127
  polynomial::MonomialBasis      include/polatory/polynomial/monomial_basis.hpp:24-330
  interpolation::Solver::solve   include/polatory/interpolation/solver.hpp:75-142 (the loop only)

The operator is the saddle-point matrix
    [ A + nugget I   F    P_mu    ]
    [ F^T            H    P_sigma ]      A, F, F^T, H = the four evaluator kinds per RBF
    [ P^T                 0       ]
applied to CUDA vectors: every FMM evaluation reads and writes HBM directly (zero-copy across
the C ABI); torch is used for the vector plumbing and the skinny polynomial products only.

Multi-GPU (`group` given, one rank per GPU): points are put in the Morton order of
the symmetric evaluator's tree, rank r owns the contiguous range `[lo, hi)` of its target shard
and the Krylov vectors are sharded accordingly (the `l` polynomial coefficients live on the last
rank).  One matvec = one all-gather of the weight shards (NCCL over NVLink), the PARTITIONED upward
pass (each rank computes the multipoles below the level-cut cells it owns or needs, one all-gather of the
level-cut expansions, the upper levels on every rank: `plt_eval_set_partition`), and the downward pass + near
field of the rank's own leaves.  Gradient data (sigma > 0) is supported on
one GPU only in this round.
"""
from __future__ import annotations

import numpy as np

from . import fmm


class Model:
    """What Operator needs from polatory::Model (include/polatory/model.hpp): the RBFs, the
    polynomial degree (-1 = none) and the nugget."""

    def __init__(self, rbfs, poly_degree=-1, nugget=0.0):
        self.rbfs = list(rbfs) if isinstance(rbfs, (list, tuple)) else [rbfs]
        self.dim = self.rbfs[0].dim
        self.poly_degree = int(poly_degree)
        self.nugget = float(nugget)

    def poly_basis_size(self):
        # polynomial_basis_base.hpp: C(dim + degree, degree)
        d, k = self.dim, self.poly_degree
        if k < 0:
            return 0
        from math import comb
        return comb(d + k, k)


def monomial_basis(dim, degree, points, grad_points=None):
    """MonomialBasis<Dim>::evaluate(points, grad_points): (mu + dim*sigma) x l matrix; value rows
    first, then `dim` derivative rows per gradient point (monomial_basis.hpp:31-330).  Column
    order: 1 | x y z | x^2 xy xz y^2 yz z^2 (the dim-restricted prefix of it)."""
    points = np.asarray(points, dtype=np.float64).reshape(-1, dim)
    gp = np.zeros((0, dim)) if grad_points is None else np.asarray(grad_points, dtype=np.float64).reshape(-1, dim)
    mu, sigma = len(points), len(gp)
    exps = [tuple([0] * dim)]
    if degree >= 1:
        exps += [tuple(1 if a == b else 0 for b in range(dim)) for a in range(dim)]
    if degree >= 2:
        for a in range(dim):
            for b in range(a, dim):
                e = [0] * dim
                e[a] += 1
                e[b] += 1
                exps.append(tuple(e))
    if degree < 0:
        exps = []
    out = np.zeros((mu + dim * sigma, len(exps)))
    for c, e in enumerate(exps):
        col = np.ones(mu)
        for a in range(dim):
            col = col * points[:, a] ** e[a]
        out[:mu, c] = col
        for k in range(dim):  # d/dx_k of the monomial at the gradient points
            if e[k] == 0:
                continue
            col = np.full(sigma, float(e[k]))
            for a in range(dim):
                p = e[a] - (1 if a == k else 0)
                col = col * gp[:, a] ** p
            out[mu + k:mu + dim * sigma:dim, c] = col
    return out


class Operator:
    """interpolation::Operator: `set_points`, `size()`, `__call__(weights)`; plus `apply(x, y)` on
    CUDA tensors for the device Krylov solver."""

    def __init__(self, model, bbox, accuracy=float("inf"), grad_accuracy=float("inf"), group=None, device=None):
        import torch
        self._torch = torch
        self.model = model
        self.dim = model.dim
        self.l = model.poly_basis_size()
        self.accuracy = accuracy
        self.grad_accuracy = grad_accuracy
        self.group = group
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.rank, self.world = 0, 1
        if group is not None:
            import torch.distributed as dist
            self._dist = dist
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        # operator.hpp:44-49
        self.a = [fmm.make_fmm_symmetric_evaluator(r, bbox) for r in model.rbfs]
        self.f = [fmm.make_fmm_gradient_evaluator(r, bbox) for r in model.rbfs]
        self.ft = [fmm.make_fmm_gradient_transpose_evaluator(r, bbox) for r in model.rbfs]
        self.h = [fmm.make_fmm_hessian_symmetric_evaluator(r, bbox) for r in model.rbfs]
        self.mu = self.sigma = 0
        self.lo, self.hi = 0, 0
        self.perm = None
        self.p = None

    # -- operator.hpp:83-107 -----------------------------------------------------------
    def set_points(self, points, grad_points=None):
        torch = self._torch
        dim = self.dim
        points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, dim)
        gp = np.zeros((0, dim)) if grad_points is None else \
            np.ascontiguousarray(grad_points, dtype=np.float64).reshape(-1, dim)
        self.mu, self.sigma = len(points), len(gp)
        n_rbf = len(self.a)
        acc = (self.accuracy / 2.0 if self.sigma > 0 else self.accuracy) / n_rbf
        gacc = (self.grad_accuracy / 2.0 if self.sigma > 0 else self.grad_accuracy) / n_rbf
        if self.world > 1:
            if self.sigma > 0:
                raise NotImplementedError("sharded operator: gradient data is single-GPU only in this round")
            # Morton order of the symmetric evaluator's tree = the order shards are cut in
            self.a[0].set_points(points)
            self.perm = self.a[0].permutation()
            points = np.ascontiguousarray(points[self.perm])
            # partition of the level-cut cells by Morton key range, balanced by point count
            from .parallel import DEFAULT_CUT_LEVEL, partition_keys
            self.cut = DEFAULT_CUT_LEVEL[dim]
            self.key_begin = partition_keys(self.a[0].point_keys(points, self.cut), self.world, dim, self.cut)
        else:
            self.perm = np.arange(self.mu)
        self._points = points
        self._grad_points = gp
        for i in range(n_rbf):
            self.a[i].set_points(points)
            self.a[i].set_accuracy(acc)
            if self.sigma > 0:
                self.f[i].set_source_points(gp)
                self.f[i].set_target_points(points)
                self.ft[i].set_source_points(points)
                self.ft[i].set_target_points(gp)
                self.h[i].set_points(gp)
                self.f[i].set_accuracy(acc)
                self.ft[i].set_accuracy(gacc)
                self.h[i].set_accuracy(gacc)
        self.lo, self.hi = 0, self.mu
        if self.world > 1:
            for i in range(n_rbf):
                self.a[i].set_partition(self.rank, self.world, self.cut, self.key_begin, self.group)
            self.lo, self.hi = self.a[0].target_shard_range()
            for i in range(1, n_rbf):
                assert self.a[i].target_shard_range() == (self.lo, self.hi)
            # shard sizes of every rank (the l polynomial coefficients live on the last rank)
            sizes = torch.zeros(self.world, dtype=torch.int64, device=self.device)
            sizes[self.rank] = (self.hi - self.lo) + (self.l if self.rank == self.world - 1 else 0)
            self._dist.all_reduce(sizes, op=self._dist.ReduceOp.SUM, group=self.group)
            self.shard_sizes = [int(v) for v in sizes.cpu()]
        if self.l > 0:
            p = monomial_basis(dim, self.model.poly_degree, points, gp)
            self.p = torch.from_numpy(p).to(self.device)
            q = p.copy()  # common::orthonormalize_cols (modified Gram-Schmidt), solver.hpp:62-66
            for i in range(self.l):
                q[:, i] /= np.linalg.norm(q[:, i])
                for j in range(i + 1, self.l):
                    q[:, j] -= (q[:, i] @ q[:, j]) * q[:, i]
            self.p_orth = torch.from_numpy(q).to(self.device)
        m = self.mu + dim * self.sigma
        self._full = torch.zeros(m + self.l, dtype=torch.float64, device=self.device)
        self._tmp_mu = torch.empty(self.mu, dtype=torch.float64, device=self.device)
        self._res_mu = torch.empty(self.mu, dtype=torch.float64, device=self.device)
        self._tmp_sg = torch.empty(dim * self.sigma, dtype=torch.float64, device=self.device)

    def size(self):
        """Global size mu + dim*sigma + l (operator.hpp:109)."""
        return self.mu + self.dim * self.sigma + self.l

    def local_size(self):
        """Length of this rank's shard of a Krylov vector."""
        if self.world == 1:
            return self.size()
        return (self.hi - self.lo) + (self.l if self.rank == self.world - 1 else 0)

    # -- shard <-> global (caller order) ---------------------------------------------------
    def scatter(self, global_vec):
        """This rank's shard of a global caller-order vector (numpy or tensor)."""
        torch = self._torch
        g = torch.as_tensor(global_vec, dtype=torch.float64).to(self.device)
        if self.world == 1:
            return g.clone()
        perm = torch.from_numpy(self.perm[self.lo:self.hi].astype(np.int64)).to(self.device)
        parts = [g[perm]]
        if self.rank == self.world - 1:
            parts.append(g[self.mu:])
        return torch.cat(parts)

    def gather(self, local_vec):
        """Global caller-order vector from the shards (one all_reduce)."""
        torch = self._torch
        if self.world == 1:
            return local_vec.clone()
        full = self._assemble(local_vec).clone()
        out = torch.empty_like(full)
        out[torch.from_numpy(self.perm.astype(np.int64)).to(self.device)] = full[:self.mu]
        out[self.mu:] = full[self.mu:]
        return out

    def _assemble(self, x_local):
        """Full (Morton-order) vector on every rank from the shards: the all-gather of the weights (the shards
        are contiguous Morton ranges in rank order, the polynomial tail on the last rank, so their
        concatenation IS the full vector)."""
        from .parallel import allgather_shards
        return allgather_shards(x_local, self.shard_sizes, self.group, out=self._full)

    # -- operator.hpp:52-81 ------------------------------------------------------------------
    def apply(self, x, y):
        mu, ds, l = self.mu, self.dim * self.sigma, self.l
        m = mu + ds
        if self.world == 1:
            full, out = x, y
        else:
            full, out = self._assemble(x), None
        w_mu, w_sg = full[:mu], full[mu:m]
        acc_mu = self._tmp_mu
        # y.head(mu) = nugget * w.head(mu) + sum_i (A_i w_mu + F_i w_sigma)
        res_mu = out[:mu] if out is not None else self._res_mu
        res_sg = out[mu:m] if out is not None else None
        for i in range(len(self.a)):
            self.a[i].set_weights(w_mu)
            if i == 0:
                self.a[i].evaluate(res_mu)
            else:
                self.a[i].evaluate(acc_mu)
                res_mu += acc_mu
            if ds:
                self.f[i].set_weights(w_sg)
                self.f[i].evaluate(acc_mu)
                res_mu += acc_mu
                self.ft[i].set_weights(w_mu)
                self.ft[i].evaluate(res_sg if i == 0 else self._tmp_sg)
                if i > 0:
                    res_sg += self._tmp_sg
                self.h[i].set_weights(w_sg)
                self.h[i].evaluate(self._tmp_sg)
                res_sg += self._tmp_sg
        if self.model.nugget != 0.0:
            if self.world == 1:
                res_mu.add_(w_mu, alpha=self.model.nugget)
            else:
                res_mu[self.lo:self.hi].add_(w_mu[self.lo:self.hi], alpha=self.model.nugget)
        if self.world == 1:
            if l:
                y[:m] += self.p @ x[m:]
                y[m:] = self.p.T @ x[:m]
            return y
        nloc = self.hi - self.lo
        y[:nloc] = res_mu[self.lo:self.hi]
        if l:
            y[:nloc] += self.p[self.lo:self.hi] @ full[m:]
            if self.rank == self.world - 1:
                y[nloc:] = self.p.T @ full[:m]
        return y

    def orthogonalize_against_polynomials(self, x):
        """solver.hpp:88-96: the RBF part of an initial solution is made orthogonal to the polynomial space,
        weights.head(m) -= P (P^T weights.head(m)) (the RAS preconditioner relies on v.tail(l) ~ 0,
        ras_preconditioner.hpp:186-187).  `x` is this rank's shard; modified in place and returned."""
        if self.l == 0:
            return x
        m = self.mu + self.dim * self.sigma
        if self.world == 1:
            x[:m] -= self.p_orth @ (self.p_orth.T @ x[:m])
            return x
        nloc = self.hi - self.lo
        p_loc = self.p_orth[self.lo:self.hi]
        dot = p_loc.T @ x[:nloc]
        self._dist.all_reduce(dot, op=self._dist.ReduceOp.SUM, group=self.group)
        x[:nloc] -= p_loc @ dot
        return x

    def __call__(self, weights):
        torch = self._torch
        x = torch.as_tensor(weights, dtype=torch.float64).to(self.device)
        y = torch.empty_like(x)
        self.apply(x, y)
        return y


class ShardedPreconditioner:
    """Right preconditioner for sharded Krylov vectors out of a preconditioner that works on the full
    (caller-order) vector and is replicated on every rank (e.g. ras.RasPreconditioner): assemble the shards
    (one all_reduce), apply, keep this rank's shard.  The FMM matvec and the Krylov reductions scale with the
    ranks; the preconditioner is computed redundantly (a domain-sharded RAS is the next step)."""

    def __init__(self, op, preconditioner):
        self.op, self.pc = op, preconditioner

    def apply(self, x_local, y_local):
        full = self.op.gather(x_local)
        y_local.copy_(self.op.scatter(self.pc(full)))
        return y_local

    __call__ = apply


class ResidualEvaluator:
    """interpolation::ResidualEvaluator (include/polatory/interpolation/residual_evaluator.hpp:26-164), value and
    Hermite data: the residual |fit - value| is first measured EXACTLY (direct sums on the device, the reference's
    DirectEvaluator) on at most 1024 sampled value points and 1024 sampled gradient points with non-zero data, and
    only when that passes on ALL points through the fast evaluator at the user's accuracy (`fast_op`, the
    reference's separate SymmetricEvaluator of solver.hpp:41; its rows [0, m) are SymmetricEvaluator::evaluate() +
    nugget * w).  The sample is the reference's own: std::shuffle with a default-seeded std::mt19937 followed by
    std::partition, run natively (`plt_residual_sample_indices`)."""

    K_DIRECT_TARGETS = 1024  # kDirectEvaluatorTargetSize

    def __init__(self, fast_op):
        self.op = fast_op
        self._direct = []

    @staticmethod
    def _sample(values_host, n, block):
        import ctypes
        from . import _lib
        lib = _lib.load()
        out = np.empty(n, dtype=np.int64)
        v = np.ascontiguousarray(values_host, dtype=np.float64)
        if n:
            st = lib.plt_residual_sample_indices(v.ctypes.data, n, block, out.ctypes.data)
            if st != _lib.PLT_OK:
                raise RuntimeError("plt_residual_sample_indices failed")
        return out

    def set_values(self, values):
        """values: CUDA tensor (mu + dim * sigma), the right-hand side without the l zeros."""
        import torch
        op = self.op
        mu, sigma, dim = op.mu, op.sigma, op.dim
        self.values = values
        host = values.detach().cpu().numpy()
        self.direct_mu = min(mu, self.K_DIRECT_TARGETS)
        self.direct_sigma = min(sigma, self.K_DIRECT_TARGETS)
        self.idx = self._sample(host[:mu], mu, 1)[:self.direct_mu]
        self.gidx = self._sample(host[mu:], sigma, dim)[:self.direct_sigma]
        dev = op.device
        self.idx_dev = torch.from_numpy(self.idx).to(dev)
        self.gidx_dev = torch.from_numpy(self.gidx).to(dev)
        self.grows_dev = (mu + dim * self.gidx_dev[:, None] + torch.arange(dim, device=dev)[None, :]).reshape(-1)
        self.exact = self.direct_mu == mu and self.direct_sigma == sigma
        pts, gp = op._points, op._grad_points
        tp = np.ascontiguousarray(pts[self.idx])
        tg = np.ascontiguousarray(gp[self.gidx])
        bbox = fmm.Bbox.from_points(np.concatenate([pts, gp]))
        # DirectEvaluator (interpolation/direct_evaluator.hpp:38-88): exact sums, the four blocks that exist
        self._direct = []
        for rbf in op.model.rbfs:
            evs = {}
            for name, make, s_pts, t_pts in (("a", fmm.make_fmm_evaluator, pts, tp),
                                             ("f", fmm.make_fmm_gradient_evaluator, gp, tp),
                                             ("ft", fmm.make_fmm_gradient_transpose_evaluator, pts, tg),
                                             ("h", fmm.make_fmm_hessian_evaluator, gp, tg)):
                if len(s_pts) == 0 or len(t_pts) == 0:
                    continue
                ev = make(rbf, bbox)
                ev.force_direct(True)
                ev.set_source_points(s_pts)
                ev.set_target_points(t_pts)
                evs[name] = ev
            self._direct.append(evs)
        self._buf_v = torch.empty(self.direct_mu, dtype=torch.float64, device=dev)
        self._buf_g = torch.empty(dim * self.direct_sigma, dtype=torch.float64, device=dev)
        self._full = torch.empty(op.size(), dtype=torch.float64, device=dev)

    def converged(self, weights, tolerance, grad_tolerance=None):
        """(converged, residual, grad_residual, exact) as residual_evaluator.hpp:55-108."""
        import torch
        op = self.op
        grad_tolerance = tolerance if grad_tolerance is None else grad_tolerance
        mu, m = op.mu, op.mu + op.dim * op.sigma
        w_mu, w_sg = weights[:mu].contiguous(), weights[mu:m].contiguous()
        fit_v = op.model.nugget * weights[self.idx_dev]
        fit_g = torch.zeros(op.dim * self.direct_sigma, dtype=torch.float64, device=op.device)
        for evs in self._direct:
            for name, w, buf in (("a", w_mu, self._buf_v), ("f", w_sg, self._buf_v),
                                 ("ft", w_mu, self._buf_g), ("h", w_sg, self._buf_g)):
                if name in evs:
                    evs[name].set_weights(w)
                    evs[name].evaluate(buf)
                    if buf is self._buf_v:
                        fit_v = fit_v + buf
                    else:
                        fit_g = fit_g + buf
        if op.l:
            fit_v = fit_v + op.p[self.idx_dev] @ weights[m:]
            if self.direct_sigma:
                fit_g = fit_g + op.p[self.grows_dev] @ weights[m:]
        res = float((fit_v - self.values[self.idx_dev]).abs().max()) if self.direct_mu else 0.0
        gres = float((fit_g - self.values[self.grows_dev]).abs().max()) if self.direct_sigma else 0.0
        if res > tolerance or gres > grad_tolerance:
            return False, res, gres, self.exact
        if self.exact:
            return True, res, gres, True
        op.apply(weights, self._full)
        res = float((self._full[:mu] - self.values[:mu]).abs().max()) if mu else 0.0
        gres = float((self._full[mu:m] - self.values[mu:m]).abs().max()) if op.sigma else 0.0
        return (res <= tolerance and gres <= grad_tolerance), res, gres, True


def solve(op, values, tolerance, max_iter, preconditioner=None, initial_weights=None, residual_op=None,
          grad_tolerance=None, verbose=False):
    """The loop of interpolation::Solver::solve (solver.hpp:75-142) on the device: FGMRES with an optional right
    preconditioner over `op` (the matvec operator -- the reference builds it at accuracy 0, solver.hpp:40);
    `values` is this rank's shard of the right-hand side (without the `l` zeros, which are appended here).
    Convergence on one GPU: the reference's ResidualEvaluator over `residual_op`, a second Operator at the user's
    accuracy (solver.hpp:41; default: `op` itself).  Sharded vectors: max-norm of the true residual through the
    operator once the solver's own 2-norm estimate allows it.
    Returns (weights shard, iteration count)."""
    import torch
    from .krylov import Fgmres
    dev = op.device
    values = torch.as_tensor(values, dtype=torch.float64).to(dev)
    n_tail = op.local_size() - values.numel()
    rhs = torch.cat([values, torch.zeros(n_tail, dtype=torch.float64, device=dev)])
    solver = Fgmres(op, rhs, max_iter, group=op.group)
    if initial_weights is not None:
        solver.set_initial_solution(op.orthogonalize_against_polynomials(
            torch.as_tensor(initial_weights, dtype=torch.float64).to(dev).clone()))
    if preconditioner is not None:
        solver.set_right_preconditioner(preconditioner)
    solver.setup()
    weights = solver.solution_vector()
    if solver.relative_residual() == 0.0:
        return weights, 0
    if op.group is None:
        # the reference's convergence test: exact residual on a sample first, then everywhere (solver.hpp:112-135)
        res_eval = ResidualEvaluator(op if residual_op is None else residual_op)
        res_eval.set_values(values)
        while True:
            weights = solver.solution_vector()
            ok, res, gres, exact = res_eval.converged(weights, tolerance, grad_tolerance)
            if verbose:
                print(f"{solver.iteration_count():8d} {'' if exact else '~'}{res:12.4e} {gres:12.4e}", flush=True)
            if ok:
                return weights, solver.iteration_count()
            if solver.iteration_count() == solver.max_iterations():
                raise RuntimeError("reached the maximum number of iterations")  # solver.hpp:132-134
            solver.iterate_process()
    n_glob = op.size()
    fit = torch.empty_like(rhs)
    while True:
        weights = solver.solution_vector()
        if solver.absolute_residual() <= tolerance * np.sqrt(n_glob):
            op.apply(weights, fit)
            res = (fit[:values.numel()] - values).abs().max() if values.numel() else torch.zeros((), device=dev)
            if op.group is not None:
                import torch.distributed as dist
                dist.all_reduce(res, op=dist.ReduceOp.MAX, group=op.group)
            if float(res) <= tolerance:
                return weights, solver.iteration_count()
        if solver.iteration_count() == solver.max_iterations():
            raise RuntimeError("reached the maximum number of iterations")  # solver.hpp:132-134
        solver.iterate_process()


class Solver:
    """interpolation::Solver (include/polatory/interpolation/solver.hpp:22-154): the matvec operator at accuracy 0
    (-> order 12, d 8: `op_(model, points, grad_points, 0.0, 0.0)`, :40), a separate residual evaluator at the
    user's accuracy (:41) and the RAS preconditioner (:60)."""

    def __init__(self, model, points, grad_points=None, accuracy=float("inf"), grad_accuracy=float("inf"),
                 device=None, ras_kwargs=None):
        from .ras import RasPreconditioner
        dim = model.dim
        points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, dim)
        gp = np.zeros((0, dim)) if grad_points is None else \
            np.ascontiguousarray(grad_points, dtype=np.float64).reshape(-1, dim)
        bbox = fmm.Bbox.from_points(np.concatenate([points, gp]))
        self.model = model
        self.op = Operator(model, bbox, 0.0, 0.0, device=device)
        self.res_op = Operator(model, bbox, accuracy, grad_accuracy, device=device)
        self.op.set_points(points, gp if len(gp) else None)
        self.res_op.set_points(points, gp if len(gp) else None)
        self.pc = RasPreconditioner(model, points, gp if len(gp) else None, device=device, **(ras_kwargs or {}))
        self.iterations = None

    def solve(self, values, tolerance, grad_tolerance=None, max_iter=100, initial_weights=None, verbose=False):
        w, it = solve(self.op, values, tolerance, max_iter, preconditioner=self.pc.apply,
                      initial_weights=initial_weights, residual_op=self.res_op, grad_tolerance=grad_tolerance,
                      verbose=verbose)
        self.iterations = it
        return w


class Fitter:
    """interpolation::Fitter (include/polatory/interpolation/fitter.hpp:12-36)."""

    def __init__(self, model, points, grad_points=None):
        self.model, self.points, self.grad_points = model, points, grad_points
        self.solver = None

    def fit(self, values, tolerance, grad_tolerance=None, max_iter=100, accuracy=float("inf"),
            grad_accuracy=float("inf"), initial_weights=None, verbose=False):
        self.solver = Solver(self.model, self.points, self.grad_points, accuracy, grad_accuracy)
        return self.solver.solve(values, tolerance, grad_tolerance, max_iter, initial_weights, verbose)
