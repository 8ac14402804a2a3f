"""GPU experiment: config C1 (benchmark/predict) -- cov_exponential dual kriging: 100k random points in the unit
cube fitted (degree 0, tolerance 1e-4, accuracy infinity), then evaluated on a 100^3 grid."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import polatory_b200 as pb
from polatory_b200.operator import Model, Operator, solve
from polatory_b200.ras import RasPreconditioner
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
rng = np.random.default_rng(0)
pts = rng.uniform(0, 1, (n, 3))
vals = np.sin(np.pi * pts).sum(axis=1)
rbf = pb.make_rbf("exp", [1.0, 0.2])
model = Model(rbf, poly_degree=0, nugget=0.0)
g = np.linspace(0, 1, 100)
grid = np.ascontiguousarray(np.stack(np.meshgrid(g, g, g, indexing="ij"), axis=-1).reshape(-1, 3))
bbox = pb.Bbox(np.zeros(3), np.ones(3))
torch.cuda.synchronize(); t0 = time.time()
op = Operator(model, bbox); op.set_points(pts)
pc = RasPreconditioner(model, pts)
torch.cuda.synchronize(); t1 = time.time()
w, iters = solve(op, vals, 1e-4, 100, preconditioner=pc.apply)
torch.cuda.synchronize(); t2 = time.time()
ev = pb.make_fmm_evaluator(rbf, bbox)
ev.set_source_points(pts); ev.set_target_points(grid); ev.set_weights(w[:n].cpu().numpy())
pred = ev.evaluate() + float(w[n])
t3 = time.time()
print(f"C1: n={n} setup {t1-t0:.2f}s fit {t2-t1:.2f}s ({iters} it, {pc.n_levels} levels) predict 1M grid {t3-t2:.3f}s total {t3-t0:.2f}s")
from oracle import fmm as ofmm
sub = rng.choice(len(grid), 300, replace=False)
ref = ofmm.direct("exp", [1.0, 0.2], 3, 0, pts, grid[sub], w[:n].cpu().numpy()) + float(w[n])
print("prediction vs exact sums: max abs", np.max(np.abs(pred[sub] - ref)), " field range", pred.min(), pred.max())
sub2 = rng.choice(n, 300, replace=False)
fit = ofmm.direct("exp", [1.0, 0.2], 3, 0, pts, pts[sub2], w[:n].cpu().numpy()) + float(w[n])
print("fit residual on exact samples:", np.max(np.abs(fit - vals[sub2])))
