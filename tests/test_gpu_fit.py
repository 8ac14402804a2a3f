"""The reference's own fit acceptance tests on the GPU path, in the reference's configuration.

  test/interpolation/test_fitter.cpp:24-73   th3 (Triharmonic3D {1.0}), random anisotropy, poly_degree = cpd_order - 1
  = 1, nugget 0.01, tolerance = grad_tolerance = 1e-3, max_iter 100, accuracy = tolerance / 100;
  shapes (10 000, 0), (10 000, 10 000) and "the special case" (1, 10 000).
  include/polatory/interpolation/solver.hpp:40-41   matvec operator at accuracy 0 (-> order 12, d 8), a separate
  residual evaluator at the user's accuracy, RAS right preconditioner, FGMRES.

Acceptance as in the reference: a SymmetricEvaluator at the user's accuracy + nugget * w reproduces the data to the
tolerance (max norm, values and gradients).  Here also: a sample of rows against exact direct sums (oracle).
"""
import numpy as np
import pytest

from conftest import random_anisotropy

pytestmark = pytest.mark.gpu


def _sample_data(n, aniso, rng):
    """test/utility.hpp:44-67 (seeded): points uniform in [-1, 1]^3, values sum_j sin(pi (A p)_j)."""
    pts = rng.uniform(-1, 1, (n, 3))
    return pts, np.sin(np.pi * (pts @ aniso.T)).sum(axis=1)


def _sample_grad_data(n, aniso, rng):
    """test/utility.hpp:69-93: gradients pi cos(pi (A p)_j) contracted with A."""
    pts = rng.uniform(-1, 1, (n, 3))
    return pts, (np.pi * np.cos(np.pi * (pts @ aniso.T))) @ aniso


@pytest.mark.parametrize("n_points,n_grad_points", [(10000, 0), (10000, 10000), (1, 10000)])
def test_reference_fitter_shapes(n_points, n_grad_points, rng):
    import torch
    import polatory_b200 as pb
    from polatory_b200.operator import Fitter, Model, Operator
    from oracle import direct as odir
    from oracle import rbf as orbf
    dim = 3
    tol = gtol = 1e-3
    acc = gacc = tol / 100.0
    aniso = random_anisotropy(dim, rng)
    pts, vals = _sample_data(n_points, aniso, rng)
    gpts, gvals = _sample_grad_data(n_grad_points, aniso, rng)
    rhs = np.concatenate([vals, gvals.reshape(-1)])
    model = Model(pb.make_rbf("th3", [1.0], dim, aniso), poly_degree=1, nugget=0.01)
    fitter = Fitter(model, pts, gpts if n_grad_points else None)
    w = fitter.fit(rhs, tol, gtol, 100, acc, gacc)
    solver = fitter.solver
    assert w.numel() == n_points + dim * n_grad_points + model.poly_basis_size()
    # the reference's configuration of the fit (solver.hpp:40): matvec at accuracy 0 -> (12, 8) on the FMM branch
    if n_points >= 1024:
        assert (solver.op.a[0].config()["order"], solver.op.a[0].config()["d"]) == (12, 8)
    assert solver.iterations <= 100
    # acceptance of test_fitter.cpp:57-70: fresh evaluator at the user's accuracy (+ nugget on the value rows)
    ev = Operator(model, pb.Bbox.from_points(np.concatenate([pts, gpts])), acc, gacc)
    ev.set_points(pts, gpts if n_grad_points else None)
    fit = ev(w)[:n_points + dim * n_grad_points].cpu().numpy()
    assert np.max(np.abs(fit[:n_points] - rhs[:n_points])) < tol
    if n_grad_points:
        assert np.max(np.abs(fit[n_points:] - rhs[n_points:])) < gtol
    # exact sums on a sample of rows
    wh = w.cpu().numpy()
    o_rbf = orbf.make_rbf("th3", [1.0], dim, aniso)
    sub = np.sort(rng.choice(n_points, min(n_points, 64), replace=False))
    gsub = np.sort(rng.choice(n_grad_points, 64, replace=False)) if n_grad_points else np.zeros(0, dtype=int)
    exact = odir.direct_evaluator(o_rbf, 0.0, pts, gpts, wh[:n_points + dim * n_grad_points], pts[sub], gpts[gsub])
    from polatory_b200.operator import monomial_basis
    exact += monomial_basis(dim, 1, pts[sub], gpts[gsub]) @ wh[-4:]
    exact[:len(sub)] += 0.01 * wh[sub]
    want = np.concatenate([vals[sub], gvals[gsub].reshape(-1)])
    assert np.max(np.abs(exact - want)) < 1.5 * tol
    print(f"fit {n_points}+{n_grad_points}: {solver.iterations} iterations, levels {solver.pc.n_levels}")


def test_config_c4_at_size():
    """BASELINE.json config #4 at full size: th3 (Triharmonic3D) with gradient constraints (Hermite-Birkhoff),
    500 000 value points + 500 000 gradient points in [-1, 1]^3 (2 000 000 rows), anisotropic model (10^(+-0.25)
    stretch x rotation), linear polynomial, in the reference's fit configuration and with the parameters of the
    reference's own th3 fit test (test/interpolation/test_fitter.cpp:27-45: nugget 0.01, tolerance 1e-3 for values
    and gradients, accuracy tolerance / 100).  4-level RAS, FGMRES; acceptance: the reference's ResidualEvaluator
    converged (exact sample + every row through the fast evaluator) and exact sums on sampled rows."""
    import polatory_b200 as pb
    from polatory_b200.operator import Model, Solver, monomial_basis
    from oracle import direct as odir
    from oracle import rbf as orbf
    n, dim, tol, nugget = 500_000, 3, 1e-3, 0.01
    pts = np.random.default_rng(0).uniform(-1, 1, (n, dim))
    gpts = np.random.default_rng(1).uniform(-1, 1, (n, dim))
    q, _ = np.linalg.qr(np.random.default_rng(2).standard_normal((dim, dim)))
    aniso = np.diag(10.0 ** np.array([0.25, 0.0, -0.25])) @ q
    values = np.concatenate([np.sin(np.pi * (pts @ aniso.T)).sum(axis=1),
                             ((np.pi * np.cos(np.pi * (gpts @ aniso.T))) @ aniso).reshape(-1)])
    model = Model(pb.make_rbf("th3", [1.0, 0.0], dim, aniso), poly_degree=1, nugget=nugget)
    solver = Solver(model, pts, gpts, tol / 100, tol / 100)
    assert solver.pc.n_levels == 4
    w = solver.solve(values, tol, tol, max_iter=130).cpu().numpy()
    assert solver.op.a[0].config() == {"tree_height": 6, "order": 12, "d": 8}
    m = n + dim * n
    rng = np.random.default_rng(9)
    sp, sg = rng.choice(n, 48, replace=False), rng.choice(n, 24, replace=False)
    o = orbf.make_rbf("th3", [1.0, 0.0], dim, aniso)
    fit = odir.direct_evaluator(o, 0.0, pts, gpts, w[:m], pts[sp], gpts[sg]) + \
        monomial_basis(dim, 1, pts[sp], gpts[sg]) @ w[m:]
    fit[:len(sp)] += nugget * w[sp]
    ref = np.concatenate([values[sp], values[n:].reshape(n, dim)[sg].reshape(-1)])
    assert np.max(np.abs(fit[:len(sp)] - ref[:len(sp)])) < tol
    assert np.max(np.abs(fit[len(sp):] - ref[len(sp):])) < tol
    print(f"C4 at size: {solver.iterations} FGMRES iterations")
