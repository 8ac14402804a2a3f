"""CPU checks of the host-side logic of the RAS preconditioner (polatory_b200/ras.py): the level
structure, the coarse-point choice and the domain decomposition restated from
include/polatory/preconditioner/{ras_preconditioner,domain_divider,domain}.hpp."""
import numpy as np

from polatory_b200 import ras


def test_std_mt19937_and_uniform_index():
    g = ras._StdMt19937()
    assert g() == 3499211612 and g() == 581869302  # std::mt19937 default seed, first two outputs
    g = ras._StdMt19937()
    # Lemire: (x * n) >> 32 when the low word is not below the threshold
    assert g.uniform_index(1000) == (3499211612 * 1000) >> 32
    assert all(0 <= ras._StdMt19937().uniform_index(n) < n for n in (1, 2, 7, 10 ** 6))


def test_level_structure_matches_reference_formula():
    # ras_preconditioner.hpp:70-75: 1M rows -> 4 levels; coarse sizes 10^(3.311 + k * 0.896)
    n_levels, counts = ras.level_structure(1_000_000)
    assert n_levels == 4
    assert counts[0] == 2048 or counts[0] == 2047  # pow() truncation
    assert 15_000 < counts[1] < 17_000 and 120_000 < counts[2] < 135_000
    assert ras.level_structure(2048)[0] == 1 and ras.level_structure(2049)[0] == 2
    assert ras.level_structure(30_000)[0] == 3


def test_round_half_to_even():
    from oracle import ras as oras
    assert [oras._round_half_to_even(x) for x in (0.5, 1.5, 2.5, 3.5, 2.4, 2.6)] == [0, 2, 2, 4, 2, 3]


def test_divide_domains_invariants():
    rng = np.random.default_rng(0)
    n = 20_000
    pts = rng.uniform(-1, 1, (n, 3))
    poly = [17, 4242, 9001, 15000]
    idcs = np.concatenate([poly, np.setdiff1d(np.arange(n), poly)])
    doms = ras.divide_domains(pts, idcs, poly)
    inner_count = np.zeros(n, dtype=int)
    for d in doms:
        assert len(d.point_indices) <= ras.K_MAX_LEAF_SIZE + len(poly)
        assert list(d.point_indices[:len(poly)]) == poly                 # poly points first (domain.hpp:39-41)
        rest = d.point_indices[len(poly):]
        assert np.all(np.diff(rest) > 0)                                   # sorted, no duplicates
        assert not set(rest) & set(poly)
        assert len(np.unique(d.point_indices)) == len(d.point_indices)
        inner_count[d.point_indices[d.inner_point]] += 1
    assert np.all(inner_count == 1)                                        # every point is inner exactly once
    # overlap: domains are larger than their inner sets
    assert sum(len(d.point_indices) for d in doms) > 1.5 * n


def test_choose_coarse_points():
    rng = np.random.default_rng(1)
    n = 5000
    pts = rng.uniform(-1, 1, (n, 3))
    poly = [3]
    idcs = np.concatenate([poly, np.setdiff1d(np.arange(n), poly)])
    c = ras.choose_coarse_points(pts, idcs, poly, 500)
    assert c[0] == 3 and len(c) == 501 and len(np.unique(c)) == 501
    # cluster centres spread over the cloud: every octant is represented
    oct_ = ((pts[c[1:]] > 0) * [1, 2, 4]).sum(axis=1)
    assert len(np.unique(oct_)) == 8


def test_unisolvent_and_lagrange():
    rng = np.random.default_rng(2)
    pts = rng.uniform(-1, 1, (1000, 3))
    assert ras.unisolvent_point_set(pts, 0, 3) == [(3499211612 * 1000) >> 32]  # degree 0: the first draw
    idx = ras.unisolvent_point_set(pts, 1, 3)
    assert len(idx) == 4 and idx == sorted(idx)
    lag = ras.lagrange_basis_matrix(pts, idx, 1, 3)
    np.testing.assert_allclose(lag[idx], np.eye(4), atol=1e-10)   # cardinal at the poly points
    np.testing.assert_allclose(lag.sum(axis=1), 1.0, atol=1e-10)  # partition of unity (degree >= 0)


def test_native_matches_oracle_restatement():
    rng = np.random.default_rng(3)
    n = 30_000
    pts = rng.uniform(-1, 1, (n, 3)) * [3.0, 1.0, 0.5]
    poly = [5, 777, 12345, 20000]
    idcs = np.concatenate([poly, np.setdiff1d(np.arange(n), poly)])
    from oracle import ras as oras
    a = ras.divide_domains(pts, idcs, poly)
    b = oras.divide_domains(pts, idcs, poly)
    key = lambda d: (tuple(d.point_indices[:8]), len(d.point_indices))
    a, b = sorted(a, key=key), sorted(b, key=key)
    assert len(a) == len(b)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x.point_indices, y.point_indices)
        np.testing.assert_array_equal(x.inner_point, y.inner_point)
    for target in (64, 1000, 3001, 9000, 20000):  # the deep targets reach two-point clusters (centre ties)
        ca = ras.choose_coarse_points(pts, idcs, poly, target)
        cb = oras.choose_coarse_points(pts, idcs, poly, target)
        assert list(ca[:4]) == poly and len(ca) == len(cb) == target + 4
        assert set(ca.tolist()) == set(cb.tolist())


def test_oracle_ras_preconditions_dense_system():
    """oracle/ras.py + oracle/krylov.py on a 2-level bh3 problem with linear polynomial: right-preconditioned
    FGMRES reaches 1e-8 in a handful of iterations and the weights are orthogonal to the polynomials."""
    from oracle.krylov import Fgmres
    from oracle.ras import RasOracle, monomials
    rng = np.random.default_rng(1)
    n = 3000
    pts = rng.uniform(-1, 1, (n, 3))
    a = -np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(axis=2))
    o = RasOracle(a, pts, 3, 1, 0.0, [5, 100, 900, 2000])
    assert o.n_levels == 2 and len(o.point_idcs[0]) in (2047 + 4, 2048 + 4)  # pow() truncation
    p = monomials(3, 1, pts)
    full = np.block([[a, p], [p.T, np.zeros((4, 4))]])
    v = np.concatenate([np.sin(np.pi * pts).sum(axis=1), np.zeros(4)])
    s = Fgmres(lambda x: full @ x, v, 30)
    s.set_right_preconditioner(o)
    s.setup()
    for _ in range(8):
        s.iterate_process()
        if s.relative_residual() < 1e-8:
            break
    assert s.relative_residual() < 1e-8
    x = s.solution_vector()
    assert np.linalg.norm(full @ x - v) <= 1e-7 * np.linalg.norm(v)
    assert np.max(np.abs(p.T @ x[:n])) <= 1e-8 * np.max(np.abs(x[:n])) * n


def test_native_mixed_matches_oracle_restatement():
    """Hermite bookkeeping (value + gradient points, multiplicity-weighted cuts): native C++ vs the oracle's numpy."""
    from oracle import ras as oras
    rng = np.random.default_rng(8)
    mu, sg = 9000, 4000
    p = rng.uniform(-1, 1, (mu, 3)) * [1.0, 2.0, 0.6]
    g = rng.uniform(-1, 1, (sg, 3)) * [1.0, 2.0, 0.6]
    poly = [11, 500, 4242, 8000]
    pid = np.concatenate([poly, np.setdiff1d(np.arange(mu), poly)])
    gid = np.arange(sg)
    a = ras.divide_domains_mixed(p, g, pid, gid, poly)
    b = oras.divide_domains_mixed(p, g, pid, gid, poly)
    assert len(a) == len(b)
    key_a = sorted((tuple(d.point_indices.tolist()), tuple(d.inner_point.tolist()),
                    tuple(sorted(zip(d.grad_point_indices.tolist(), d.inner_grad_point.tolist())))) for d in a)
    key_b = sorted((tuple(pi.tolist()), tuple(pin.tolist()), tuple(sorted(zip(gi.tolist(), gin.tolist()))))
                   for pi, pin, gi, gin in b)
    assert key_a == key_b
    own_p, own_g = np.zeros(mu, int), np.zeros(sg, int)
    for d in a:
        assert len(d.point_indices) + 3 * len(d.grad_point_indices) <= ras.K_MAX_LEAF_SIZE + len(poly)
        own_p[d.point_indices[d.inner_point]] += 1
        own_g[d.grad_point_indices[d.inner_grad_point]] += 1
    assert np.all(own_p == 1) and np.all(own_g == 1)
    for target in (300, 2000, 5001):
        pa, ga = ras.choose_coarse_points_mixed(p, g, pid, gid, poly, target)
        pb_, gb = oras.choose_coarse_points_mixed(p, g, pid, gid, poly, target)
        assert list(pa[:4]) == poly and set(pa.tolist()) == set(pb_.tolist()) and set(ga.tolist()) == set(gb.tolist())
        assert target <= (len(pa) - 4) + 3 * len(ga) < target + 3


def test_oracle_ras_hermite_preconditions_dense_system():
    """oracle/ras.py on Hermite data (values + gradients, anisotropic th3, linear polynomial, 2 levels): the
    right-preconditioned FGMRES of oracle/krylov.py reaches 1e-10 in a handful of iterations on the exact dense
    saddle-point system, and the interpolation conditions hold for values and gradients."""
    from oracle import direct as odir, rbf as orbf
    from oracle.krylov import Fgmres
    from oracle.ras import RasOracle, monomials
    rng = np.random.default_rng(41)
    dim, mu, sigma = 3, 1500, 300
    aniso = np.diag([1.5, 1.0, 0.7])
    pts, gpts = rng.uniform(-1, 1, (mu, dim)), rng.uniform(-1, 1, (sigma, dim))
    m = mu + dim * sigma
    o_rbf = orbf.make_rbf("th3", [1.0, 0.0], dim, aniso)
    from conftest import dense_th3_hermite
    a = dense_th3_hermite(pts, gpts, aniso)     # closed-form mat_a, spot-checked against the exact direct evaluator
    for col in (5, mu + 4, m - 1):
        ref = odir.direct_evaluator(o_rbf, 0.0, pts, gpts, np.eye(m)[:, col], pts, gpts)
        assert np.max(np.abs(a[:, col] - ref)) <= 1e-11 * np.max(np.abs(ref))
    assert np.max(np.abs(a - a.T)) <= 1e-12 * np.max(np.abs(a))
    o = RasOracle(a, pts, dim, 1, 0.0, [3, 400, 900, 1400], grad_points=gpts, a_points=pts @ aniso.T,
                  a_grad_points=gpts @ aniso.T)
    assert o.n_levels == 2
    own = np.zeros(m, dtype=int)
    for g in o.fine[1]:
        own[g["idx"][g["inner"]]] += 1
    assert np.all(own == 1)   # every row (value or gradient component) is owned by exactly one fine domain
    p = monomials(dim, 1, pts, gpts)
    full = np.block([[a, p], [p.T, np.zeros((4, 4))]])
    values = np.concatenate([np.sin(np.pi * pts).sum(axis=1), (np.pi * np.cos(np.pi * gpts)).reshape(-1)])
    v = np.concatenate([values, np.zeros(4)])
    s = Fgmres(lambda x: full @ x, v, 30)
    s.set_right_preconditioner(o)
    s.setup()
    for _ in range(12):
        s.iterate_process()
        if s.relative_residual() < 1e-10:
            break
    assert s.relative_residual() < 1e-10 and s.iteration_count() <= 12
    x = s.solution_vector()
    assert np.max(np.abs((full @ x)[:m] - values)) <= 1e-7
