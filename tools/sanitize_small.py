"""Small end-to-end run for compute-sanitizer: evaluation (FMM branch, all four kinds), sharded target range, fit with RAS
(value and Hermite data)."""
import sys, numpy as np, torch
sys.path.insert(0, '.')
import polatory_b200 as pb
from polatory_b200.operator import Model, Operator, solve
from polatory_b200.ras import RasPreconditioner
rng = np.random.default_rng(0)
src = rng.uniform(-1, 1, (6000, 3)); trg = rng.uniform(-1, 1, (5000, 3))
bbox = pb.Bbox(-np.ones(3), np.ones(3))
for kind, make in enumerate((pb.make_fmm_evaluator, pb.make_fmm_gradient_evaluator,
                             pb.make_fmm_gradient_transpose_evaluator, pb.make_fmm_hessian_evaluator)):
    ev = make(pb.make_rbf("th3", [1.0, 0.0]), bbox)
    ev.set_source_points(src); ev.set_target_points(trg)
    ev.set_weights(rng.uniform(-1, 1, 6000 * ev.km)); ev.evaluate()
    ev.set_target_shard(1, 3); ev.evaluate()
pts = rng.uniform(-1, 1, (3000, 3)); gp = rng.uniform(-1, 1, (400, 3))
vals = np.sin(pts).sum(axis=1)
m = Model(pb.make_rbf("bh3", [1.0, 0.0]), poly_degree=0)
op = Operator(m, bbox, accuracy=1e-7); op.set_points(pts)
w, it = solve(op, vals, 1e-5, 50, preconditioner=RasPreconditioner(m, pts).apply)
print("value fit iterations", it)
m2 = Model(pb.make_rbf("th3", [1.0, 0.0]), poly_degree=1)
op2 = Operator(m2, bbox, accuracy=1e-6, grad_accuracy=1e-6); op2.set_points(pts, gp)
v2 = np.concatenate([vals, np.cos(gp).reshape(-1)])
w2, it2 = solve(op2, v2, 1e-4, 80, preconditioner=RasPreconditioner(m2, pts, gp).apply)
print("hermite fit iterations", it2)
