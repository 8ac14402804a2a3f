"""The oracle's RBF formulas, pinned the way the reference pins its own:
test/rbf/test_rbf.cpp:47-174 (anisotropy identity, analytic gradient / Hessian vs central
finite differences, h = 1e-8, tolerance 1e-4)."""
import numpy as np
import pytest

from conftest import ALL_RBFS, default_params, random_anisotropy
from oracle import direct as odir
from oracle import fmm as ofmm
from oracle import rbf as orbf


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("name", ALL_RBFS)
def test_anisotropy_identity(name, dim, rng):
    # test_rbf.cpp:28-45: rbf.evaluate(x) == iso_rbf.evaluate(A x)
    a = random_anisotropy(dim, rng)
    r = orbf.make_rbf(name, default_params(name), dim, a)
    iso = orbf.make_rbf(name, default_params(name), dim)
    x = rng.uniform(-1, 1, (20, dim))
    np.testing.assert_allclose(r.evaluate(x), iso.evaluate(x @ a.T), rtol=1e-14, atol=1e-15)


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("name", ALL_RBFS)
def test_gradient_finite_difference(name, dim, rng):
    # test_rbf.cpp:47-79
    h, tol = 1e-8, 1e-4
    a = random_anisotropy(dim, rng)
    r = orbf.make_rbf(name, default_params(name), dim, a)
    xs = rng.uniform(-1, 1, (10, dim))
    g = r.evaluate_gradient(xs)
    for i in range(dim):
        e = np.zeros(dim)
        e[i] = h
        approx = (r.evaluate(xs + e) - r.evaluate(xs - e)) / (2 * h)
        np.testing.assert_allclose(g[:, i], approx, atol=tol, rtol=tol)


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("name", [n for n in ALL_RBFS if n not in ("sph", "cub")])
def test_hessian_finite_difference(name, dim, rng):
    # test_rbf.cpp:81-122 (skipped for cub / sph as in :163-170)
    h, tol = 1e-8, 1e-4
    a = random_anisotropy(dim, rng)
    r = orbf.make_rbf(name, default_params(name), dim, a)
    xs = rng.uniform(-1, 1, (10, dim))
    hess = r.evaluate_hessian(xs)
    for i in range(dim):
        e = np.zeros(dim)
        e[i] = h
        approx = (r.evaluate_gradient(xs + e) - r.evaluate_gradient(xs - e)) / (2 * h)
        np.testing.assert_allclose(hess[:, i, :], approx, atol=tol, rtol=tol)


def test_hessian_of_compact_rbfs_throws():
    # cov_spherical.hpp:53-55, cov_cubic.hpp:58-60
    for name in ("sph", "cub"):
        with pytest.raises(RuntimeError):
            orbf.make_rbf(name, [1.0, 1.0], 3).evaluate_hessian(np.zeros((1, 3)))


@pytest.mark.parametrize("name", ["sp3", "sp5", "sp7", "sp9"])
def test_spheroidal_split_sums_to_full(name, rng):
    # src/fmm/spheroidal_evaluator.hpp:24-29: full = direct part + fast part
    r = orbf.make_rbf(name, [1.1, 0.7], 3)
    x = rng.uniform(-1, 1, (200, 3))
    np.testing.assert_allclose(r.direct_part().evaluate(x) + r.fast_part().evaluate(x), r.evaluate(x),
                               rtol=1e-13, atol=1e-15)
    assert np.isfinite(r.direct_part().support_radius_isotropic())
    assert not np.isfinite(r.fast_part().support_radius_isotropic())


def test_singular_point_conventions():
    # polyharmonic gradient/Hessian are 0 at rho == 0; bh2/th2 value is 0.
    z = np.zeros((1, 3))
    for name in ("bh3", "th3", "bh2", "th2"):
        r = orbf.make_rbf(name, [1.0, 0.0], 3)
        assert np.all(r.evaluate_gradient(z) == 0.0)
        assert np.all(r.evaluate_hessian(z) == 0.0)
        assert r.evaluate(z)[0] == 0.0


@pytest.mark.parametrize("dim", [1, 2, 3])
@pytest.mark.parametrize("name", ALL_RBFS)
def test_c_oracle_matches_numpy_oracle(name, dim, rng):
    """Two independent restatements (numpy, C) of the same reference formulas agree."""
    a = random_anisotropy(dim, rng)
    params = default_params(name)
    for kind in range(4):
        if kind == 3 and name in ("sph", "cub"):
            continue
        for part in ((0, 1, 2) if name in ("sp3", "sp5", "sp7", "sp9") else (0,)):
            for sym in (False, True):
                if sym and kind in (1, 2):
                    continue
                src = rng.uniform(-1, 1, (40, dim))
                trg = rng.uniform(-1, 1, (30, dim))
                w = rng.uniform(-1, 1, 40 * odir.kind_km(kind, dim))
                o = orbf.make_rbf(name, params, dim, a, part)
                with np.errstate(all="ignore"):
                    ref = odir.full_direct(o, kind, src, trg, w, symmetric=sym)
                got = ofmm.direct(name, params, dim, kind, src, trg, w, a, part, sym)
                if np.isnan(ref).any():
                    assert (np.isnan(ref) == np.isnan(got)).all()
                    continue
                scale = max(np.max(np.abs(ref)), 1e-300)
                tol = 1e-10 if (dim == 1 and kind == 3) else 1e-12  # 1-D Hessians cancel to ~0
                assert np.max(np.abs(ref - got)) / scale < tol
