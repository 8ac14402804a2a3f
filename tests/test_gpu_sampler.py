"""GPU tests of the bulk Evaluator (interpolation::Evaluator) and of the isosurface field sampler pattern
(isosurface::RbfFieldFunction driven per lattice layer, include/polatory/isosurface/rmt/lattice.hpp:421-445)."""
import numpy as np
import pytest

from conftest import random_anisotropy

pytestmark = pytest.mark.gpu


def test_evaluator_matches_direct_evaluator(rng):
    """test/interpolation/test_evaluator.cpp:27-75 through the composed Evaluator: th3, random anisotropy, nugget is
    irrelevant for evaluation, 1024 points + 1024 gradient points -> 1024 + 1024 gradient targets, polynomial degree
    1, accuracy 1e-4 for values and gradients, against the exact DirectEvaluator (oracle)."""
    import polatory_b200 as pb
    from oracle import direct as odir
    from oracle import rbf as orbf
    from polatory_b200.evaluator import Evaluator
    from polatory_b200.operator import Model, monomial_basis
    dim, n, acc = 3, 1024, 1e-4
    a = random_anisotropy(dim, rng)
    pts, gpts, epts, gepts = (rng.uniform(-1, 1, (n, dim)) for _ in range(4))
    model = Model(pb.make_rbf("th3", [1.0], dim, a), poly_degree=1, nugget=0.01)
    w = rng.uniform(-1, 1, n + dim * n + model.poly_basis_size())
    ev = Evaluator(model, pts, gpts, pb.Bbox(-np.ones(dim), np.ones(dim)), acc, acc)
    ev.set_weights(w)
    got = ev.evaluate(epts, gepts).cpu().numpy()
    ref = odir.direct_evaluator(orbf.make_rbf("th3", [1.0], dim, a), 0.0, pts, gpts, w[:n + dim * n], epts, gepts) + \
        monomial_basis(dim, 1, epts, gepts) @ w[n + dim * n:]
    assert ev.a[0].config()["tree_height"] == 3
    assert np.max(np.abs(got[:n] - ref[:n])) < acc
    assert np.max(np.abs(got[n:] - ref[n:])) < acc


def test_field_sampler_batches_reuse_the_upward_pass():
    """The isosurface pattern: fixed centres + weights, one batch of lattice nodes per layer.  The batches equal the
    one-shot evaluation bit for bit and only the first one runs P2M / M2M / the multipole transforms."""
    import torch
    import polatory_b200 as pb
    from polatory_b200 import workloads as wl
    from polatory_b200.evaluator import Evaluator, RbfFieldFunction
    from polatory_b200.operator import Model
    src, _ = wl.sdf_offset_cloud(210_000, seed=3)
    rng = np.random.default_rng(5)
    model = Model(pb.make_rbf("bh3", [1.0, 0.0]), poly_degree=0)
    w = np.concatenate([rng.uniform(-1, 1, len(src)), [0.3]])
    lo, hi = 1.1 * src.min(axis=0), 1.1 * src.max(axis=0)
    shape = (40, 70, 70)
    grid = wl.grid_points(lo, hi, shape)
    field = RbfFieldFunction(model, src, w)
    field.set_evaluation_bbox(pb.Bbox(lo, hi))
    per_layer = shape[1] * shape[2]
    d_grid = torch.from_numpy(grid).cuda()
    out, upward_runs = [], 0
    for layer in range(shape[0]):
        out.append(field(d_grid[layer * per_layer:(layer + 1) * per_layer]).clone())
        upward_runs += "p2m" in field.evaluator.a[0].phase_times()
        assert field.evaluator.a[0].config()["tree_height"] == 6
    batched = torch.cat(out).cpu().numpy()
    assert field.batches == shape[0] and upward_runs == 1
    one_shot = Evaluator(model, src, None, pb.Bbox(lo, hi).convex_hull(pb.Bbox.from_points(src)))
    one_shot.set_weights(w)
    assert pb.fmm.tree_height(3, max(len(src), len(grid))) == 6
    ref = one_shot.evaluate(d_grid).cpu().numpy()
    assert np.array_equal(batched, ref)
    # and it is the interpolant: exact sums on a few nodes
    from oracle import fmm as ofmm
    sub = rng.choice(len(grid), 100, replace=False)
    exact = ofmm.direct("bh3", [1.0, 0.0], 3, 0, src, grid[sub], w[:-1]) + 0.3
    assert np.max(np.abs(batched[sub] - exact)) < 2e-5 * np.max(np.abs(exact))
