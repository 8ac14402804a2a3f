"""GPU tests of the device FGMRES (csrc/krylov.cu) and the RBF operator (polatory_b200/operator.py)
against the oracle (numpy restatements of src/krylov/*.cpp and the exact direct sums).

Tolerances: the Krylov iterates of the device solver and of the oracle follow the same arithmetic
up to summation order -- residual histories agree to 1e-10 relative, iteration counts exactly on
these problems (north_star: +-1); operator vs dense direct matrix: the evaluator tolerances of
tests/test_gpu_parity.py.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from conftest import random_anisotropy
from test_oracle_krylov import _reference_problem

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def torch():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.mark.parametrize("with_x0", [False, True])
@pytest.mark.parametrize("with_pc", [False, True])
def test_fgmres_reference_acceptance_and_oracle(torch, with_x0, with_pc):
    """test/krylov/test_krylov.cpp:84-110 on the device solver, plus iterate-by-iterate agreement
    with the oracle."""
    from oracle.krylov import Fgmres as OracleFgmres
    from polatory_b200.krylov import Fgmres
    m, pc, rhs, x0 = _reference_problem()
    n = len(rhs)
    dev = torch.device("cuda")
    tm, tpc = torch.from_numpy(m).to(dev), torch.from_numpy(pc).to(dev)
    s = Fgmres(lambda x, y: torch.mv(tm, x, out=y), torch.from_numpy(rhs).to(dev), n)
    o = OracleFgmres(lambda v: m @ v, rhs, n)
    if with_x0:
        s.set_initial_solution(torch.from_numpy(x0).to(dev))
        o.set_initial_solution(x0)
    if with_pc:
        s.set_right_preconditioner(lambda x, y: torch.mv(tpc, x, out=y))
        o.set_right_preconditioner(lambda v: pc @ v)
    with pytest.raises(RuntimeError):
        s.set_left_preconditioner(None)  # fgmres.hpp:15-17
    s.setup()
    o.setup()
    assert abs(s.relative_residual() - o.relative_residual()) <= 1e-12 * max(1.0, o.relative_residual())
    last = 0.0
    for i in range(s.max_iterations()):
        s.iterate_process()
        o.iterate_process()
        x = s.solution_vector().cpu().numpy()
        cur = s.relative_residual()
        assert abs(np.linalg.norm(rhs - m @ x) / np.linalg.norm(rhs) - cur) < 1e-12
        if i > 0:
            assert cur < last
        last = cur
        if o.relative_residual() > 1e-9:  # above the rounding floor the two histories coincide
            assert abs(cur - o.relative_residual()) <= 1e-8 * o.relative_residual()
    assert s.iteration_count() == n == o.iteration_count()
    s.iterate_process()  # at max_iter: no-op (gmres.cpp:10-12)
    assert s.iteration_count() == n
    assert s.launch_count() > 0


def _dense_operator(orbf, odir, rbf_name, params, dim, aniso, pts, gpts, degree, nugget):
    """The saddle-point matrix of operator.hpp:52-81 from exact direct sums (column by column)."""
    from polatory_b200.operator import monomial_basis
    o = orbf.make_rbf(rbf_name, params, dim, aniso)
    mu, sigma = len(pts), len(gpts)
    m = mu + dim * sigma
    a = np.zeros((m, m))
    eye = np.eye(m)
    for c in range(m):
        col = odir.direct_evaluator(o, 0.0, pts, gpts, eye[:, c], pts, gpts)
        a[:, c] = col
    a[:mu, :mu] += nugget * np.eye(mu)
    p = monomial_basis(dim, degree, pts, gpts)
    l = p.shape[1]
    full = np.zeros((m + l, m + l))
    full[:m, :m] = a
    full[:m, m:] = p
    full[m:, :m] = p.T
    return full


@pytest.mark.parametrize("case", ["bh3_values", "th3_hermite_aniso", "bh2_2d"])
def test_operator_matches_dense_direct(torch, case):
    import polatory_b200 as pb
    from oracle import direct as odir, rbf as orbf
    from polatory_b200.operator import Model, Operator
    rng = np.random.default_rng(3)
    if case == "bh3_values":
        name, params, dim, degree, nugget, mu, sigma, aniso = "bh3", [1.0, 0.0], 3, 0, 0.01, 300, 0, np.eye(3)
    elif case == "th3_hermite_aniso":
        name, params, dim, degree, nugget, mu, sigma = "th3", [1.0, 0.0], 3, 1, 0.0, 200, 60
        aniso = random_anisotropy(3, rng)
    else:
        name, params, dim, degree, nugget, mu, sigma, aniso = "bh2", [1.0, 0.0], 2, 1, 0.0, 250, 0, np.eye(2)
    pts = rng.uniform(-1, 1, (mu, dim))
    gpts = rng.uniform(-1, 1, (sigma, dim))
    dense = _dense_operator(orbf, odir, name, params, dim, aniso, pts, gpts, degree, nugget)
    model = Model(pb.make_rbf(name, params, dim, aniso), poly_degree=degree, nugget=nugget)
    op = Operator(model, pb.Bbox(-np.ones(dim), np.ones(dim)))
    op.set_points(pts, gpts)
    assert op.size() == dense.shape[0]
    w = rng.uniform(-1, 1, op.size())
    got = op(w).cpu().numpy()
    ref = dense @ w
    assert np.max(np.abs(got - ref)) <= 1e-11 * np.max(np.abs(ref))


def test_fit_small_bh3_matches_oracle_fgmres(torch):
    """A complete (unpreconditioned) fit: device FGMRES over the FMM-branch operator against the
    oracle FGMRES over the exact dense matrix -- same iteration count (+-1), weights to the solver
    tolerance, interpolation conditions met."""
    import polatory_b200 as pb
    from oracle import direct as odir, rbf as orbf
    from oracle.krylov import Fgmres as OracleFgmres
    from polatory_b200.operator import Model, Operator, solve
    rng = np.random.default_rng(11)
    mu, dim = 1500, 3
    pts = rng.uniform(-1, 1, (mu, dim))
    values = np.sin(np.pi * pts).sum(axis=1)
    nugget = 0.05  # keeps the unpreconditioned system well enough conditioned for a short test
    # the saddle-point matrix in closed form (bh3: phi = -r, polyharmonic_odd.hpp:32-45; degree 0: one column of ones)
    a = -np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(axis=2)) + nugget * np.eye(mu)
    dense = np.block([[a, np.ones((mu, 1))], [np.ones((1, mu)), np.zeros((1, 1))]])
    model = Model(pb.make_rbf("bh3", [1.0, 0.0]), poly_degree=0, nugget=nugget)
    op = Operator(model, pb.Bbox(-np.ones(dim), np.ones(dim)), accuracy=0.0)  # order 12 / d 8
    op.set_points(pts)
    tol, max_iter = 1e-6, 400
    w, iters = solve(op, values, tol, max_iter)
    assert op.a[0].config()["tree_height"] > 0  # the FMM branch
    w = w.cpu().numpy()
    # oracle: same loop over the exact matrix
    rhs = np.concatenate([values, [0.0]])
    o = OracleFgmres(lambda v: dense @ v, rhs, max_iter)
    o.setup()
    o_iters = None
    while True:
        x = o.solution_vector()
        if o.absolute_residual() <= tol * np.sqrt(len(rhs)) and np.max(np.abs((dense @ x)[:mu] - values)) <= tol:
            o_iters = o.iteration_count()
            break
        o.iterate_process()
    assert abs(iters - o_iters) <= 1, (iters, o_iters)
    assert np.max(np.abs((dense @ w)[:mu] - values)) <= 2 * tol
    assert np.max(np.abs(w - x)) <= 1e-4 * np.max(np.abs(x))


_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["PLT_ROOT"])
import polatory_b200 as pb
from polatory_b200.operator import Model, Operator, solve
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(0)
dist.init_process_group("gloo", rank=rank, world_size=world)  # two ranks share cuda:0 in this test
rng = np.random.default_rng(5)
mu, dim = 6000, 3
pts = rng.uniform(-1, 1, (mu, dim))
values = np.sin(np.pi * pts).sum(axis=1)
model = Model(pb.make_rbf("bh3", [1.0, 0.0]), poly_degree=0, nugget=0.05)
bbox = pb.Bbox(-np.ones(dim), np.ones(dim))
# evaluator accuracy two orders below the fit tolerance, as the reference's fitter sets it
single = Operator(model, bbox, accuracy=1e-7); single.set_points(pts)
sharded = Operator(model, bbox, accuracy=1e-7, group=dist.group.WORLD); sharded.set_points(pts)
assert sharded.a[0].config is not None
w = rng.uniform(-1, 1, single.size())
ref = single(w)
loc = sharded.scatter(w)
assert loc.numel() == sharded.local_size()
y = torch.empty_like(loc); sharded.apply(loc, y)
got = sharded.gather(y)
err = float((got - ref).abs().max() / ref.abs().max())
assert err < 1e-12, err
sizes = torch.tensor([sharded.local_size()], dtype=torch.int64); dist.all_reduce(sizes)
assert int(sizes) == single.size()
w1, it1 = solve(single, values, 1e-5, 300)
wl, it2 = solve(sharded, sharded.scatter(np.concatenate([values, [0.0]]))[:sharded.hi - sharded.lo], 1e-5, 300)
w2 = sharded.gather(wl)
assert abs(it1 - it2) <= 1, (it1, it2)
assert float((w1 - w2).abs().max() / w1.abs().max()) < 1e-6
# preconditioned fit: replicated RAS behind the sharded matvec (operator.ShardedPreconditioner)
from polatory_b200.operator import ShardedPreconditioner
from polatory_b200.ras import RasPreconditioner
model0 = Model(pb.make_rbf("bh3", [1.0, 0.0]), poly_degree=0, nugget=0.0)
s1 = Operator(model0, bbox, accuracy=1e-7); s1.set_points(pts)
sN = Operator(model0, bbox, accuracy=1e-7, group=dist.group.WORLD); sN.set_points(pts)
pc = RasPreconditioner(model0, pts)
wa, ia = solve(s1, values, 1e-5, 60, preconditioner=pc.apply)
wb, ib = solve(sN, sN.scatter(np.concatenate([values, [0.0]]))[:sN.hi - sN.lo], 1e-5, 60,
               preconditioner=ShardedPreconditioner(sN, pc).apply)
assert abs(ia - ib) <= 1 and ia <= 15, (ia, ib)
assert float((wa - sN.gather(wb)).abs().max() / wa.abs().max()) < 1e-5
if rank == 0:
    print("sharded ok", err, it1, it2, ia, ib)
dist.destroy_process_group()
"""


def test_sharded_operator_and_fgmres_two_ranks(torch, tmp_path):
    """World size 2 (gloo, both ranks on cuda:0): the sharded matvec reassembles to the single-rank
    matvec to 1e-12 and the sharded FGMRES reproduces the single-rank fit (iterations +-1)."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   PLT_ROOT=ROOT)
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    assert "sharded ok" in outs[0]
