"""Which objects of a fit sit in reference cycles (they keep device memory alive until a gc pass)?"""
import gc, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import polatory_b200 as pb
from polatory_b200.operator import Model, Operator, ResidualEvaluator, solve
from polatory_b200.ras import RasPreconditioner
rng = np.random.default_rng(0)
pts = rng.uniform(-1, 1, (6000, 3)); vals = np.sin(pts).sum(axis=1)
model = Model(pb.make_rbf("bh3", [1.0, 0.0]), poly_degree=0, nugget=0.0)
bbox = pb.Bbox(-np.ones(3), np.ones(3))
gc.collect(); gc.disable()
op = Operator(model, bbox, 0.0, 0.0); res_op = Operator(model, bbox, 1e-6, 1e-6)
op.set_points(pts); res_op.set_points(pts)
pc = RasPreconditioner(model, pts)
w, it = solve(op, vals, 1e-4, 50, preconditioner=pc.apply, residual_op=res_op)
print("iterations", it)
del op, res_op, pc, w
gc.set_debug(gc.DEBUG_SAVEALL)
n = gc.collect()
import collections
print("unreachable:", n, collections.Counter(type(o).__name__ for o in gc.garbage).most_common(12))
for o in gc.garbage:
    if type(o).__name__ in ("RasPreconditioner", "Operator", "Fgmres", "FmmGenericEvaluator", "_FineLevel"):
        print("  in a cycle:", type(o).__name__, [type(r).__name__ for r in gc.get_referrers(o)][:6])
