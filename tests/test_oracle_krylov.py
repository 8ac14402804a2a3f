"""CPU checks of the Krylov oracle (oracle/krylov.py) against the reference's own acceptance test,
and of the host-side polynomial basis."""
import numpy as np
import pytest

from oracle.krylov import Fgmres


def _reference_problem(n=100, seed=0):
    """test/krylov/test_krylov.cpp:22-75: random symmetric operator, perturbed-inverse preconditioner."""
    rng = np.random.default_rng(seed)
    m = (rng.uniform(-1, 1, (n, n)) + 1.0) / 2.0
    m = np.triu(m) + np.triu(m, 1).T
    pc = np.linalg.inv(m) + 0.1 * np.diag(rng.uniform(-1, 1, n))
    solution = (rng.uniform(-1, 1, n) + 1.0) / 2.0
    return m, pc, m @ solution, rng.uniform(-1, 1, n)


@pytest.mark.parametrize("with_x0", [False, True])
@pytest.mark.parametrize("with_pc", [False, True])
def test_fgmres_oracle_reference_acceptance(with_x0, with_pc):
    """test_krylov.cpp:84-110 (TEST_F(KrylovTest, fgmres)): the solver's residual estimate equals
    the true relative residual to 1e-12 after every iteration and decreases monotonically."""
    m, pc, rhs, x0 = _reference_problem()
    s = Fgmres(lambda v: m @ v, rhs, len(rhs))
    if with_x0:
        s.set_initial_solution(x0)
    if with_pc:
        s.set_right_preconditioner(lambda v: pc @ v)
    s.setup()
    last = 0.0
    for i in range(s.max_iterations()):
        s.iterate_process()
        x = s.solution_vector()
        cur = s.relative_residual()
        assert abs(np.linalg.norm(rhs - m @ x) / np.linalg.norm(rhs) - cur) < 1e-12
        if i > 0:
            assert cur < last
        last = cur
    assert s.iteration_count() == len(rhs)


def test_monomial_basis_layout():
    """monomial_basis.hpp: 3-D degree 2 columns 1 x y z x^2 xy xz y^2 yz z^2; gradient rows
    mu + 3 i + k hold d/dx_k."""
    from polatory_b200.operator import Model, monomial_basis
    import polatory_b200 as pb
    p = np.array([[2.0, 3.0, 5.0]])
    g = np.array([[7.0, 11.0, 13.0]])
    m = monomial_basis(3, 2, p, g)
    assert m.shape == (4, 10)
    np.testing.assert_array_equal(m[0], [1, 2, 3, 5, 4, 6, 10, 9, 15, 25])
    np.testing.assert_array_equal(m[1], [0, 1, 0, 0, 14, 11, 13, 0, 0, 0])   # d/dx
    np.testing.assert_array_equal(m[2], [0, 0, 1, 0, 0, 7, 0, 22, 13, 0])    # d/dy
    np.testing.assert_array_equal(m[3], [0, 0, 0, 1, 0, 0, 7, 0, 11, 26])    # d/dz
    m2 = monomial_basis(2, 1, np.array([[2.0, 3.0]]), np.array([[5.0, 7.0]]))
    np.testing.assert_array_equal(m2, [[1, 2, 3], [0, 1, 0], [0, 0, 1]])
    assert monomial_basis(3, -1, p).shape == (1, 0)
    assert Model(pb.make_rbf("bh3", [1.0]), poly_degree=1).poly_basis_size() == 4
    assert Model(pb.make_rbf("bh2", [1.0], 2), poly_degree=2).poly_basis_size() == 6
    assert Model(pb.make_rbf("bh3", [1.0]), poly_degree=-1).poly_basis_size() == 0
