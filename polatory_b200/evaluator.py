"""Host-side mirror of the reference's bulk evaluator and of the isosurface field sampler over the device evaluators.

  interpolation::Evaluator<Dim>     include/polatory/interpolation/evaluator.hpp:20-173
  isosurface::RbfFieldFunction      include/polatory/isosurface/rbf_field_function.hpp:12-34
  Interpolant::evaluate_impl / set_evaluation_bbox_impl   include/polatory/interpolant.hpp:71-75,201-210
  the caller pattern                include/polatory/isosurface/rmt/lattice.hpp:421-445 (one batch of lattice nodes per
                                    layer / wavefront), isosurface/vertex_refiner.hpp:67-77 (4 samples per vertex)

`Evaluator` composes the four generic evaluator kinds per RBF plus the polynomial, with the reference's vector
layouts (weights [mu values | dim * sigma gradient weights | l polynomial coefficients], result [values at the target
points | dim gradient components at the target gradient points]).

What the device path adds for the sampler's many-batches pattern: the reference frees both trees at
the end of every evaluate() (src/fmm/fmm_evaluator.hpp:107-109) and therefore redoes the source tree, P2M, M2M and the
multipole transforms for every batch; here the source tree, the sorted weights and the multipole spectra stay
resident in HBM for as long as the sources and weights do not change, so a batch costs the target-side work only
(target tree, interaction plan, M2L, L2L, L2P, P2P).  The tree height follows src/fmm/utility.hpp:12-16 per batch
(max(n_src, n_trg)), so batches up to the number of sources all share one height and one cached upward pass.
"""
from __future__ import annotations

import numpy as np

from . import fmm
from .operator import monomial_basis


class Evaluator:
    """interpolation::Evaluator(model, source_points, source_grad_points, bbox, accuracy, grad_accuracy)."""

    def __init__(self, model, source_points=None, source_grad_points=None, bbox=None, accuracy=float("inf"),
                 grad_accuracy=float("inf"), device=None):
        import torch
        self._torch = torch
        self.model = model
        self.dim = model.dim
        self.l = model.poly_basis_size()
        self.accuracy, self.grad_accuracy = accuracy, grad_accuracy
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        dim = self.dim
        sp = None if source_points is None else np.ascontiguousarray(source_points, dtype=np.float64).reshape(-1, dim)
        sg = np.zeros((0, dim)) if source_grad_points is None else \
            np.ascontiguousarray(source_grad_points, dtype=np.float64).reshape(-1, dim)
        if bbox is None:
            if sp is None:
                raise ValueError("a bbox is required when no source points are given")
            bbox = fmm.Bbox.from_points(np.concatenate([sp, sg]))   # evaluator.hpp:33-38
        self.bbox = bbox
        # evaluator.hpp:53-58
        self.a = [fmm.make_fmm_evaluator(r, bbox) for r in model.rbfs]
        self.f = [fmm.make_fmm_gradient_evaluator(r, bbox) for r in model.rbfs]
        self.ft = [fmm.make_fmm_gradient_transpose_evaluator(r, bbox) for r in model.rbfs]
        self.h = [fmm.make_fmm_hessian_evaluator(r, bbox) for r in model.rbfs]
        self.mu = self.sigma = self.trg_mu = self.trg_sigma = 0
        self._coeffs = None
        self._p_trg = None
        if sp is not None:
            self.set_source_points(sp, sg)

    # -- evaluator.hpp:91-110 ----------------------------------------------------------------
    def set_source_points(self, points, grad_points=None):
        dim = self.dim
        points = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, dim)
        gp = np.zeros((0, dim)) if grad_points is None else \
            np.ascontiguousarray(grad_points, dtype=np.float64).reshape(-1, dim)
        self.mu, self.sigma = len(points), len(gp)
        n = len(self.a)
        acc = (self.accuracy / 2.0 if self.sigma > 0 else self.accuracy) / n
        gacc = (self.grad_accuracy / 2.0 if self.sigma > 0 else self.grad_accuracy) / n
        for i in range(n):
            self.a[i].set_source_points(points)
            self.f[i].set_source_points(gp)
            self.ft[i].set_source_points(points)
            self.h[i].set_source_points(gp)
            self.a[i].set_accuracy(acc)
            self.f[i].set_accuracy(acc)
            self.ft[i].set_accuracy(gacc)
            self.h[i].set_accuracy(gacc)

    # -- evaluator.hpp:114-128 ---------------------------------------------------------------
    def set_target_points(self, points, grad_points=None):
        """points / grad_points: numpy arrays or CUDA tensors (device-resident batches cross the ABI zero-copy)."""
        torch = self._torch
        dim = self.dim

        def shape(x):
            if x is None:
                return None, 0
            if hasattr(x, "is_cuda"):
                x = x.reshape(-1, dim)
                return x, int(x.shape[0])
            x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, dim)
            return x, len(x)

        pts, self.trg_mu = shape(points)
        gp, self.trg_sigma = shape(grad_points)
        if gp is None:
            gp = np.zeros((0, dim))
        for i in range(len(self.a)):
            self.a[i].set_target_points(pts)
            self.f[i].set_target_points(pts)
            self.ft[i].set_target_points(gp)
            self.h[i].set_target_points(gp)
        if self.l > 0:
            if hasattr(pts, "is_cuda") and self.trg_sigma == 0 and self.model.poly_degree <= 1:
                # monomials 1 | x y z on the device (polynomial/monomial_basis.hpp:31-60)
                cols = [torch.ones(self.trg_mu, 1, dtype=torch.float64, device=pts.device)]
                if self.model.poly_degree == 1:
                    cols.append(pts)
                self._p_trg = torch.cat(cols, dim=1)
            else:
                host = lambda x: x.detach().cpu().numpy() if hasattr(x, "is_cuda") else x  # noqa: E731
                self._p_trg = torch.from_numpy(monomial_basis(dim, self.model.poly_degree, host(pts), host(gp))).to(self.device)

    # -- evaluator.hpp:130-144 ---------------------------------------------------------------
    def set_weights(self, weights):
        torch = self._torch
        w = torch.as_tensor(weights, dtype=torch.float64).to(self.device)
        mu, ds = self.mu, self.dim * self.sigma
        assert w.numel() == mu + ds + self.l
        w_mu, w_sg = w[:mu].contiguous(), w[mu:mu + ds].contiguous()
        for i in range(len(self.a)):
            self.a[i].set_weights(w_mu)
            self.f[i].set_weights(w_sg)
            self.ft[i].set_weights(w_mu)
            self.h[i].set_weights(w_sg)
        self._coeffs = w[mu + ds:].clone() if self.l else None

    # -- evaluator.hpp:65-89 -----------------------------------------------------------------
    def evaluate(self, target_points=None, target_grad_points=None, out=None):
        """Returns a CUDA tensor [trg_mu values | dim * trg_sigma gradient components]."""
        torch = self._torch
        if target_points is not None:
            self.set_target_points(target_points, target_grad_points)
        tm, ts = self.trg_mu, self.dim * self.trg_sigma
        y = torch.empty(tm + ts, dtype=torch.float64, device=self.device) if out is None else out
        tmp_v = torch.empty(tm, dtype=torch.float64, device=self.device) if (self.sigma or len(self.a) > 1) else None
        tmp_g = torch.empty(ts, dtype=torch.float64, device=self.device) if ts else None
        yv, yg = y[:tm], y[tm:]
        for i in range(len(self.a)):
            if i == 0:
                self.a[i].evaluate(yv)
            else:
                self.a[i].evaluate(tmp_v)
                yv += tmp_v
            if self.sigma:
                self.f[i].evaluate(tmp_v)
                yv += tmp_v
            if ts:
                if i == 0:
                    self.ft[i].evaluate(yg)
                else:
                    self.ft[i].evaluate(tmp_g)
                    yg += tmp_g
                if self.sigma:
                    self.h[i].evaluate(tmp_g)
                    yg += tmp_g
        if self.l > 0:
            y += self._p_trg @ self._coeffs
        return y

    def launch_count(self):
        return sum(e.launch_count() for e in self.a + self.f + self.ft + self.h)

    def phase_times(self):
        out = {}
        for e in self.a + self.f + self.ft + self.h:
            for k, v in e.phase_times().items():
                out[k] = out.get(k, 0.0) + v
        return out


class RbfFieldFunction:
    """isosurface::RbfFieldFunction over a fitted interpolant (model, centres, weights): `set_evaluation_bbox(bbox)`
    builds the evaluator over bbox U bbox(centres) and sets sources + weights ONCE
    (Interpolant::set_evaluation_bbox_impl, interpolant.hpp:201-210); `__call__(points)` is
    Interpolant::evaluate_impl = Evaluator::evaluate(points) for one batch of lattice nodes."""

    def __init__(self, model, centers, weights, grad_centers=None, accuracy=float("inf"), grad_accuracy=float("inf")):
        self.model = model
        self.centers = np.ascontiguousarray(centers, dtype=np.float64).reshape(-1, model.dim)
        self.grad_centers = np.zeros((0, model.dim)) if grad_centers is None else \
            np.ascontiguousarray(grad_centers, dtype=np.float64).reshape(-1, model.dim)
        self.weights = weights
        self.accuracy, self.grad_accuracy = accuracy, grad_accuracy
        self.evaluator = None
        self.batches = 0

    def set_evaluation_bbox(self, bbox):
        own = fmm.Bbox.from_points(np.concatenate([self.centers, self.grad_centers]))
        self.evaluator = Evaluator(self.model, self.centers, self.grad_centers, bbox.convex_hull(own), self.accuracy,
                                   self.grad_accuracy)
        self.evaluator.set_weights(self.weights)

    def __call__(self, points):
        if self.evaluator is None:
            raise RuntimeError("set_evaluation_bbox must be called first")
        self.batches += 1
        return self.evaluator.evaluate(points)
