"""GPU tests of the RAS preconditioner (polatory_b200/ras.py) and of the preconditioned fit.

Acceptance: the criterion of the reference's fit tests (test/interpolation/test_fitter.cpp:57-64,
test_rbf_incremental_fitter.cpp): interpolation conditions met to the absolute tolerance; plus, against
the exact dense system (oracle), the fitted weights and the Gram kernel."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


def test_gram_batched_matches_oracle(torch):
    import polatory_b200 as pb
    from conftest import random_anisotropy
    from oracle import direct as odir, rbf as orbf
    rng = np.random.default_rng(0)
    dev = torch.device("cuda")
    for name, params, dim in (("bh3", [1.0, 0.0], 3), ("exp", [1.1, 0.7], 3), ("bh2", [1.0, 0.0], 2)):
        aniso = random_anisotropy(dim, rng)
        rbf = pb.make_rbf(name, params, dim, aniso)
        ev = pb.make_fmm_evaluator(rbf, pb.Bbox(-np.ones(dim), np.ones(dim)))
        b, m = 3, 50
        pts = rng.uniform(-1, 1, (b, m, dim))
        counts = np.array([50, 37, 1], dtype=np.int32)
        out = torch.empty((b, m, m), dtype=torch.float64, device=dev)
        ev.gram_batched(torch.from_numpy(pts).to(dev), torch.from_numpy(counts).to(dev), 0.25, out)
        got = out.cpu().numpy()
        o = orbf.make_rbf(name, params, dim, aniso)
        for k in range(b):
            c = counts[k]
            ref = np.eye(m)
            cols = np.stack([odir.full_direct(o, 0, pts[k, :c], pts[k, :c], np.eye(c)[:, j]) for j in range(c)], axis=1)
            ref[:c, :c] = cols + 0.25 * np.eye(c)
            assert np.max(np.abs(got[k] - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref)))


@pytest.mark.parametrize("n,degree", [(1500, 0), (6000, 0), (6000, 1)])
def test_ras_fit_bh3(torch, n, degree):
    """FGMRES + RAS on a bh3 cloud: 1 level (coarse grid only) and 2 levels (fine domains + coarse)."""
    import polatory_b200 as pb
    from polatory_b200.operator import Model, Operator, solve
    from polatory_b200.ras import RasPreconditioner, level_structure
    rng = np.random.default_rng(7)
    pts = rng.uniform(-1, 1, (n, 3))
    values = np.sin(np.pi * pts).sum(axis=1)
    model = Model(pb.make_rbf("bh3", [1.0, 0.0]), poly_degree=degree, nugget=0.0)
    op = Operator(model, pb.Bbox(-np.ones(3), np.ones(3)), accuracy=1e-8)
    op.set_points(pts)
    pc = RasPreconditioner(model, pts)
    assert pc.n_levels == level_structure(n)[0] == (1 if n <= 2048 else 2)
    tol = 1e-6
    w, iters = solve(op, values, tol, 60, preconditioner=pc.apply)
    assert iters <= (2 if pc.n_levels == 1 else 25), iters
    # interpolation conditions against the EXACT operator (direct sums)
    from oracle import direct as odir, rbf as orbf
    from polatory_b200.operator import monomial_basis
    o = orbf.make_rbf("bh3", [1.0, 0.0], 3, np.eye(3))
    wv = w.cpu().numpy()
    sub = rng.choice(n, 200, replace=False)
    fit = odir.full_direct(o, 0, pts, pts[sub], wv[:n]) + monomial_basis(3, degree, pts[sub]) @ wv[n:]
    assert np.max(np.abs(fit - values[sub])) <= 10 * tol
    # orthogonality of the RBF weights to the polynomials (solver.hpp / ras orthogonalize)
    assert np.max(np.abs(monomial_basis(3, degree, pts).T @ wv[:n])) <= 1e-6 * np.max(np.abs(wv[:n])) * n ** 0.5


@pytest.mark.parametrize("name,params,dim,degree", [("exp", [1.0, 0.3], 3, 0), ("th3", [1.0, 0.0], 3, 1),
                                                     ("bh2", [1.0, 0.0], 2, 1)])
def test_ras_fit_other_kernels(torch, name, params, dim, degree):
    """Two-level RAS + FGMRES for a covariance kernel, a cpd-order-2 kernel with linear polynomial and the
    2-D path; acceptance as above (interpolation conditions against exact sums)."""
    import polatory_b200 as pb
    from oracle import direct as odir, rbf as orbf
    from polatory_b200.operator import Model, Operator, monomial_basis, solve
    from polatory_b200.ras import RasPreconditioner
    rng = np.random.default_rng(21)
    n = 5000
    pts = rng.uniform(-1, 1, (n, dim))
    values = np.sin(np.pi * pts).sum(axis=1)
    model = Model(pb.make_rbf(name, params, dim), poly_degree=degree, nugget=0.0)
    op = Operator(model, pb.Bbox(-np.ones(dim), np.ones(dim)), accuracy=1e-9)
    op.set_points(pts)
    pc = RasPreconditioner(model, pts)
    assert pc.n_levels == 2
    tol = 1e-6
    w, iters = solve(op, values, tol, 80, preconditioner=pc.apply)
    assert iters <= 40, iters
    o = orbf.make_rbf(name, params, dim, np.eye(dim))
    wv = w.cpu().numpy()
    sub = rng.choice(n, 200, replace=False)
    fit = odir.full_direct(o, 0, pts, pts[sub], wv[:n]) + monomial_basis(dim, degree, pts[sub]) @ wv[n:]
    assert np.max(np.abs(fit - values[sub])) <= 10 * tol


@pytest.mark.parametrize("degree", [1])
def test_ras_iteration_parity_with_oracle(torch, degree):
    """The device RAS + FGMRES against the dense numpy restatement (oracle/ras.py + oracle/krylov.py) on the same
    3-level problem: same level sizes, one application agrees to 1e-6 relative (FMM level transfers at order 6
    vs exact sums), FGMRES iteration counts within +-1 (north_star's Krylov criterion)."""
    import polatory_b200 as pb
    from oracle import direct as odir, rbf as orbf
    from oracle.krylov import Fgmres as OracleFgmres
    from oracle.ras import RasOracle, monomials
    from polatory_b200.operator import Model, Operator, solve
    from polatory_b200.ras import RasPreconditioner
    rng = np.random.default_rng(33)
    n, dim = 20_600, 3
    pts = rng.uniform(-1, 1, (n, dim))
    values = np.sin(np.pi * pts).sum(axis=1)
    model = Model(pb.make_rbf("bh3", [1.0, 0.0]), poly_degree=degree, nugget=0.0)
    op = Operator(model, pb.Bbox(-np.ones(dim), np.ones(dim)), accuracy=1e-9)
    op.set_points(pts)
    pc = RasPreconditioner(model, pts)
    assert pc.n_levels == 3

    def kernel(x, y):  # bh3, s = 1, c = 0 (polyharmonic_odd.hpp:32-45), row-chunked exact differences
        out = np.empty((len(x), len(y)))
        for i0 in range(0, len(x), 512):
            d = x[i0:i0 + 512, None, :] - y[None, :, :]
            out[i0:i0 + 512] = -np.sqrt((d * d).sum(axis=2))
        return out

    a_dense = kernel(pts, pts)
    o = RasOracle(a_dense, pts, dim, degree, 0.0, pc.poly_idcs)
    assert [len(p) for p in o.point_idcs] == [len(p) for p in pc.point_idcs]
    for lvl in range(pc.n_levels):
        assert set(o.point_idcs[lvl].tolist()) == set(np.asarray(pc.point_idcs[lvl]).tolist())
    l = pc.l
    v = np.concatenate([values, np.zeros(l)])
    ref = o(v)
    got = pc(torch.from_numpy(v).cuda()).cpu().numpy()
    # order-6 level transfers (the reference's default) carry the FMM discretisation error of that order ...
    assert np.max(np.abs(got - ref)) <= 5e-3 * np.max(np.abs(ref))
    # ... with order-12 transfers the device sweep reproduces the dense restatement
    pc12 = RasPreconditioner(model, pts, transfer_config=(12, 8))
    got12 = pc12(torch.from_numpy(v).cuda()).cpu().numpy()
    assert np.max(np.abs(got12 - ref)) <= 1e-6 * np.max(np.abs(ref))
    del pc12
    tol = 1e-6
    w, iters = solve(op, values, tol, 60, preconditioner=pc.apply)
    p = monomials(dim, degree, pts)
    full = np.block([[a_dense, p], [p.T, np.zeros((l, l))]])
    s = OracleFgmres(lambda x: full @ x, v, 60)
    s.set_right_preconditioner(o)
    s.setup()
    while True:
        x = s.solution_vector()
        if s.absolute_residual() <= tol * np.sqrt(len(v)) and np.max(np.abs((full @ x)[:n] - values)) <= tol:
            break
        s.iterate_process()
    assert abs(iters - s.iteration_count()) <= 1, (iters, s.iteration_count())
    assert np.max(np.abs(w.cpu().numpy() - x)) <= 1e-3 * np.max(np.abs(x))


def test_gram_mixed_matches_oracle(torch):
    """The Hermite mat_a (value + gradient rows, include/polatory/preconditioner/mat_a.hpp:10-61) against the
    oracle's exact direct evaluator applied to unit vectors (the four kernel kinds with anisotropy)."""
    import polatory_b200 as pb
    from conftest import random_anisotropy
    from oracle import direct as odir, rbf as orbf
    rng = np.random.default_rng(4)
    dev = torch.device("cuda")
    for name, params, dim in (("th3", [1.0, 0.0], 3), ("gau", [1.1, 0.7], 3), ("th2", [1.0, 0.1], 2)):
        aniso = random_anisotropy(dim, rng)
        ev = pb.make_fmm_evaluator(pb.make_rbf(name, params, dim, aniso), pb.Bbox(-np.ones(dim), np.ones(dim)))
        mu, sigma, pad = 14, 9, 5
        p = rng.uniform(-1, 1, (mu, dim))
        g = rng.uniform(-1, 1, (sigma, dim))
        m = mu + dim * sigma
        rows = np.concatenate([p, np.repeat(g, dim, axis=0), np.zeros((pad, dim))])
        types = np.concatenate([np.zeros(mu), np.tile(1 + np.arange(dim), sigma), -np.ones(pad)]).astype(np.int8)
        out = torch.empty((1, m + pad, m + pad), dtype=torch.float64, device=dev)
        ev.gram_mixed(torch.from_numpy(rows[None]).to(dev), torch.from_numpy(types[None]).to(dev), 0.125, out)
        got = out[0].cpu().numpy()
        o = orbf.make_rbf(name, params, dim, aniso)
        ref = np.eye(m + pad)
        for c in range(m):
            ref[:m, c] = odir.direct_evaluator(o, 0.0, p, g, np.eye(m)[:, c], p, g)
        ref[:mu, :mu] += 0.125 * np.eye(mu)
        assert np.max(np.abs(got - ref)) <= 1e-11 * max(1.0, np.max(np.abs(ref))), name
        assert np.max(np.abs(got - got.T)) <= 1e-12 * max(1.0, np.max(np.abs(got)))


def test_ras_fit_hermite_th3_aniso(torch):
    """Config C4 in small: th3, anisotropic, value + gradient data (Hermite-Birkhoff), linear polynomial; two-level
    RAS with mixed domains.  Acceptance: values and gradients reproduced to the tolerance against exact sums."""
    import polatory_b200 as pb
    from conftest import random_anisotropy
    from oracle import direct as odir, rbf as orbf
    from polatory_b200.operator import Model, Operator, monomial_basis, solve
    from polatory_b200.ras import RasPreconditioner
    rng = np.random.default_rng(17)
    dim, mu, sigma = 3, 3000, 1200
    aniso = random_anisotropy(dim, rng)
    pts = rng.uniform(-1, 1, (mu, dim))
    gpts = rng.uniform(-1, 1, (sigma, dim))
    f = lambda x: np.sin(np.pi * x).sum(axis=1)
    values = np.concatenate([f(pts), (np.pi * np.cos(np.pi * gpts)).reshape(-1)])
    model = Model(pb.make_rbf("th3", [1.0, 0.0], dim, aniso), poly_degree=1, nugget=0.0)
    bbox = pb.Bbox(-np.ones(dim), np.ones(dim))
    op = Operator(model, bbox, accuracy=1e-6, grad_accuracy=1e-6)  # two orders below the tolerance
    op.set_points(pts, gpts)
    pc = RasPreconditioner(model, pts, gpts)
    assert pc.n_levels == 2 and pc.m_rows == mu + dim * sigma
    # every row is owned (inner) by exactly one fine domain
    owned = torch.zeros(pc.m_rows, device="cuda")
    owned[pc.fine[1].inner_glob] += 1
    assert bool((owned == 1).all())
    tol = 1e-4
    w, iters = solve(op, values, tol, 100, preconditioner=pc.apply)
    assert iters <= 50, iters
    wv = w.cpu().numpy()
    o = orbf.make_rbf("th3", [1.0, 0.0], dim, aniso)
    sp, sg = rng.choice(mu, 100, replace=False), rng.choice(sigma, 50, replace=False)
    m = mu + dim * sigma
    fit = odir.direct_evaluator(o, 0.0, pts, gpts, wv[:m], pts[sp], gpts[sg]) + \
        monomial_basis(dim, 1, pts[sp], gpts[sg]) @ wv[m:]
    ref = np.concatenate([values[sp], values[mu:].reshape(sigma, dim)[sg].reshape(-1)])
    assert np.max(np.abs(fit - ref)) <= 10 * tol


def test_ras_hermite_parity_with_oracle(torch):
    """Hermite data: the device RAS (mixed domains, Hermite Gram kernel, four-kind transfers) against the dense oracle
    on the same 2-level th3 problem with anisotropy: same level sets, one sweep to 1e-8 with exact level transfers (1e-4 at order 12), FGMRES
    iteration counts within +-1."""
    import polatory_b200 as pb
    from conftest import random_anisotropy
    from oracle import direct as odir, rbf as orbf
    from oracle.krylov import Fgmres as OracleFgmres
    from oracle.ras import RasOracle, monomials
    from polatory_b200.operator import Model, Operator, solve
    from polatory_b200.ras import RasPreconditioner
    rng = np.random.default_rng(41)
    dim, mu, sigma = 3, 2600, 900
    aniso = random_anisotropy(dim, rng)
    pts, gpts = rng.uniform(-1, 1, (mu, dim)), rng.uniform(-1, 1, (sigma, dim))
    m = mu + dim * sigma
    from conftest import dense_th3_hermite
    a_dense = dense_th3_hermite(pts, gpts, aniso)
    o_rbf = orbf.make_rbf("th3", [1.0, 0.0], dim, aniso)
    for col in (3, mu + 7, m - 1):   # the closed-form builder against the oracle's direct evaluator
        ref = odir.direct_evaluator(o_rbf, 0.0, pts, gpts, np.eye(m)[:, col], pts, gpts)
        assert np.max(np.abs(a_dense[:, col] - ref)) <= 1e-11 * np.max(np.abs(ref))
    values = np.concatenate([np.sin(np.pi * pts).sum(axis=1), (np.pi * np.cos(np.pi * gpts)).reshape(-1)])
    model = Model(pb.make_rbf("th3", [1.0, 0.0], dim, aniso), poly_degree=1, nugget=0.0)
    pc = RasPreconditioner(model, pts, gpts, transfer_config="direct")   # exact level transfers
    o = RasOracle(a_dense, pts, dim, 1, 0.0, pc.poly_idcs, grad_points=gpts, a_points=pts @ aniso.T,
                  a_grad_points=gpts @ aniso.T)
    assert pc.n_levels == o.n_levels == 2
    for lvl in range(2):
        assert set(o.point_idcs[lvl].tolist()) == set(np.asarray(pc.point_idcs[lvl]).tolist())
        assert set(o.grad_idcs[lvl].tolist()) == set(np.asarray(pc.grad_idcs[lvl]).tolist())
    l = pc.l
    v = np.concatenate([values, np.zeros(l)])
    ref = o(v)
    got = pc(torch.from_numpy(v).cuda()).cpu().numpy()
    err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
    assert err <= 1e-8, err   # with exact transfers the device sweep IS the dense restatement
    pc12 = RasPreconditioner(model, pts, gpts, transfer_config=(12, 8))
    err12 = np.max(np.abs(pc12(torch.from_numpy(v).cuda()).cpu().numpy() - ref)) / np.max(np.abs(ref))
    assert err12 <= 1e-4, err12  # order-12 FMM transfers of the th3 gradient / Hessian kernels (measured 1.7e-5)
    pc = pc12
    # FGMRES over the exact dense system with the oracle RAS vs the device solver with the device RAS
    p = monomials(dim, 1, pts, gpts)
    full = np.block([[a_dense, p], [p.T, np.zeros((l, l))]])
    tol = 1e-5
    s = OracleFgmres(lambda x: full @ x, v, 80)
    s.set_right_preconditioner(o)
    s.setup()
    while True:
        x = s.solution_vector()
        if s.absolute_residual() <= tol * np.sqrt(len(v)) and np.max(np.abs((full @ x)[:m] - values)) <= tol:
            break
        s.iterate_process()
    op = Operator(model, pb.Bbox(-np.ones(dim), np.ones(dim)), accuracy=0.0, grad_accuracy=0.0)  # order 12 / d 8
    op.set_points(pts, gpts)
    w, iters = solve(op, values, tol, 80, preconditioner=pc.apply)
    assert abs(iters - s.iteration_count()) <= 1, (iters, s.iteration_count())
    assert np.max(np.abs(w.cpu().numpy() - x)) <= 1e-3 * np.max(np.abs(x))


def test_ras_nugget_parity_with_oracle(torch):
    """nugget != 0: mat_a carries the nugget on the value rows (mat_a.hpp:27-29) while `ap_` comes from
    SymmetricEvaluator and the level transfers from Evaluator, neither of which applies it
    (ras_preconditioner.hpp:165-180,287-321).  One sweep of the device RAS against the dense restatement."""
    import polatory_b200 as pb
    from oracle.ras import RasOracle
    from polatory_b200.operator import Model
    from polatory_b200.ras import RasPreconditioner
    rng = np.random.default_rng(5)
    n, dim, degree, nugget = 6000, 3, 1, 0.05
    pts = rng.uniform(-1, 1, (n, dim))
    values = np.sin(np.pi * pts).sum(axis=1)
    model = Model(pb.make_rbf("bh3", [1.0, 0.0]), poly_degree=degree, nugget=nugget)
    pc = RasPreconditioner(model, pts, transfer_config=(12, 8))
    assert pc.n_levels == 2
    d = pts[:, None, :] - pts[None, :, :]
    a_dense = -np.sqrt((d * d).sum(axis=2))
    o = RasOracle(a_dense, pts, dim, degree, nugget, pc.poly_idcs)
    v = np.concatenate([values, np.zeros(pc.l)])
    ref = o(v)
    got = pc(torch.from_numpy(v).cuda()).cpu().numpy()
    assert np.max(np.abs(got - ref)) <= 1e-6 * np.max(np.abs(ref))


def test_fit_two_rbfs(torch):
    """Several RBFs per model (operator.hpp:63-73 loops over them; mat_a sums them, mat_a.hpp:23-55): a nested
    covariance model exp + gau with different anisotropies, two-level RAS, acceptance against exact sums."""
    import polatory_b200 as pb
    from conftest import random_anisotropy
    from oracle import direct as odir, rbf as orbf
    from polatory_b200.operator import Fitter, Model
    rng = np.random.default_rng(9)
    n, dim = 5000, 3
    a1, a2 = random_anisotropy(dim, rng), random_anisotropy(dim, rng)
    pts = rng.uniform(-1, 1, (n, dim))
    values = np.sin(np.pi * pts).sum(axis=1)
    model = Model([pb.make_rbf("exp", [0.7, 0.4], dim, a1), pb.make_rbf("gau", [0.3, 0.25], dim, a2)],
                  poly_degree=0, nugget=0.01)
    tol = 1e-5
    fitter = Fitter(model, pts)
    w = fitter.fit(values, tol, max_iter=80, accuracy=tol / 100).cpu().numpy()
    assert fitter.solver.pc.n_levels == 2
    sub = rng.choice(n, 200, replace=False)
    fit = odir.full_direct(orbf.make_rbf("exp", [0.7, 0.4], dim, a1), 0, pts, pts[sub], w[:n]) + \
        odir.full_direct(orbf.make_rbf("gau", [0.3, 0.25], dim, a2), 0, pts, pts[sub], w[:n]) + \
        0.01 * w[sub] + w[n]
    assert np.max(np.abs(fit - values[sub])) <= 2 * tol


@pytest.mark.parametrize("n,l,batch", [(70, 0, 5), (200, 4, 3), (333, 1, 2), (1000, 4, 2)])
def test_dense_local_kernels_match_numpy(torch, n, l, batch):
    """csrc/ras_dense.cu against numpy: Q^T A Q (fine_grid.hpp:71-81), the batched Cholesky factor and the local solve
    lambda = Q (Q^T A Q)^-1 Q^T d (fine_grid.hpp:112-133) for symmetric positive definite blocks."""
    from polatory_b200.ras import chol_batched, chol_solve_batched, reduce_q
    rng = np.random.default_rng(n + l)
    m = n + l
    g = rng.standard_normal((batch, m, m))
    a = g @ g.transpose(0, 2, 1) / m + 2.0 * np.eye(m)
    q_top = rng.standard_normal((batch, l, n))
    vals = rng.standard_normal((batch, m))
    dev = torch.device("cuda")
    d_a = torch.from_numpy(a).to(dev)
    d_q = torch.from_numpy(q_top).to(dev) if l else None
    fac = torch.empty((batch, n, n), dtype=torch.float64, device=dev)
    if l:
        reduce_q(d_a, d_q, fac)
    else:
        fac.copy_(d_a)
    qm = np.concatenate([q_top, np.broadcast_to(np.eye(n), (batch, n, n))], axis=1)   # Q = [Q_top; I]  (m x n)
    red = qm.transpose(0, 2, 1) @ a @ qm
    assert np.max(np.abs(fac.cpu().numpy() - red)) <= 1e-12 * np.max(np.abs(red))
    info = torch.zeros(batch, dtype=torch.int32, device=dev)
    chol_batched(fac, info)
    assert int(info.abs().max()) == 0
    # factor format: L below the 32 x 32 diagonal blocks (mirrored above them), the diagonal blocks hold L11^-1
    got, ref_l = fac.cpu().numpy(), np.linalg.cholesky(red)
    for k0 in range(0, n, 32):
        k1 = min(n, k0 + 32)
        if k1 < n:
            assert np.max(np.abs(got[:, k1:, k0:k1] - ref_l[:, k1:, k0:k1])) <= 1e-11 * np.max(np.abs(ref_l))
            assert np.max(np.abs(got[:, k0:k1, k1:] - ref_l[:, k1:, k0:k1].transpose(0, 2, 1))) <= 1e-11 * np.max(np.abs(ref_l))
        inv_blk = np.linalg.inv(ref_l[:, k0:k1, k0:k1])
        assert np.max(np.abs(np.tril(got[:, k0:k1, k0:k1]) - inv_blk)) <= 1e-10 * np.max(np.abs(inv_blk))
    lam = torch.empty((batch, m), dtype=torch.float64, device=dev)
    chol_solve_batched(fac, d_q, torch.from_numpy(vals).to(dev), lam)
    ref = np.stack([qm[b] @ np.linalg.solve(red[b], qm[b].T @ vals[b]) for b in range(batch)])
    assert np.max(np.abs(lam.cpu().numpy() - ref)) <= 1e-10 * np.max(np.abs(ref))
    # one factor, many right-hand sides (the coarse grid's explicit inverse) and the matrix-vector kernel
    from polatory_b200.ras import chol_inverse, gemv
    inv = chol_inverse(fac[0])
    ref_inv = np.linalg.inv(red[0])
    assert np.max(np.abs(inv.cpu().numpy() - ref_inv)) <= 1e-10 * np.max(np.abs(ref_inv))
    xv = rng.standard_normal(n)
    assert np.max(np.abs(gemv(inv, torch.from_numpy(xv).to(dev)).cpu().numpy() - ref_inv @ xv)) <= 1e-10 * np.max(np.abs(ref_inv @ xv))
    # a matrix that is not positive definite is reported, not silently factorised
    bad = torch.from_numpy(a[:1, l:, l:].copy()).to(dev)
    bad[0, n // 2, n // 2] = -1.0
    chol_batched(bad, info[:1])
    assert int(info[0]) == n // 2 + 1


@pytest.mark.parametrize("case", ["bh3_values_3_levels", "th3_hermite_aniso", "two_rbfs", "no_polynomial", "one_level",
                                  "special_case"])
def test_native_sweep_matches_torch_sweep(torch, case):
    """RasPreconditioner::operator() behind the C ABI (plt_ras_sweep_apply, csrc/ras_sweep.cu) against the torch
    re-expression of the same sweep (`apply_reference`): same level structure, same kernels for the solves and the
    level transfers, so the two agree to rounding of the small reductions."""
    import polatory_b200 as pb
    from conftest import random_anisotropy
    from polatory_b200.operator import Model
    from polatory_b200.ras import RasPreconditioner
    rng = np.random.default_rng(23)
    dim = 3
    gpts = None
    if case == "bh3_values_3_levels":
        pts = rng.uniform(-1, 1, (32000, dim))
        model = Model(pb.make_rbf("bh3", [1.0, 0.0]), poly_degree=1, nugget=0.0)
        levels = 3
    elif case == "th3_hermite_aniso":
        pts, gpts = rng.uniform(-1, 1, (3000, dim)), rng.uniform(-1, 1, (1200, dim))
        model = Model(pb.make_rbf("th3", [1.0, 0.0], dim, random_anisotropy(dim, rng)), poly_degree=1, nugget=0.01)
        levels = 2
    elif case == "two_rbfs":
        pts = rng.uniform(-1, 1, (5000, dim))
        model = Model([pb.make_rbf("exp", [0.7, 0.4], dim, random_anisotropy(dim, rng)),
                       pb.make_rbf("gau", [0.3, 0.25], dim, random_anisotropy(dim, rng))], poly_degree=0, nugget=0.01)
        levels = 2
    elif case == "no_polynomial":
        pts = rng.uniform(-1, 1, (5000, dim))
        model = Model(pb.make_rbf("exp", [1.0, 0.3]), poly_degree=-1, nugget=0.0)
        levels = 2
    elif case == "one_level":
        pts = rng.uniform(-1, 1, (900, dim))
        model = Model(pb.make_rbf("bh3", [1.0, 0.0]), poly_degree=0, nugget=0.0)
        levels = 1
    else:  # one value point + gradient points, linear polynomial (ras_preconditioner.hpp:81-86)
        pts, gpts = rng.uniform(-1, 1, (1, dim)), rng.uniform(-1, 1, (1500, dim))
        model = Model(pb.make_rbf("th3", [1.0, 0.0]), poly_degree=1, nugget=0.01)
        levels = 2
    pc = RasPreconditioner(model, pts, gpts)
    assert pc.n_levels == levels and pc._sweep
    v = torch.from_numpy(rng.uniform(-1, 1, pc.m_rows + pc.l)).cuda()
    v[pc.m_rows:] = 0.0
    got = pc.apply(v, torch.empty_like(v))
    ref = pc.apply_reference(v, torch.empty_like(v))
    scale = float(ref.abs().max())
    assert scale > 0 and bool(torch.isfinite(got).all())
    # (the small reductions are summed in another order; the local solves amplify that rounding by their condition
    # number: 2e-11 on the th3 Hermite case, < 1e-12 on value data)
    assert float((got - ref).abs().max()) <= 1e-9 * scale, float((got - ref).abs().max()) / scale
    again = pc.apply(v, torch.empty_like(v))          # deterministic: fixed-order reductions, no atomics
    assert torch.equal(again, got)
    assert pc._sweep_lib.plt_ras_sweep_launch_count(pc._sweep) > 0
