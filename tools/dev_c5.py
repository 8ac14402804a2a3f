"""GPU experiment: config C5 without the inequality constraints -- biharmonic2d terrain: n uniform points in
[0,1]^2, z = smooth terrain, degree 1, FGMRES + RAS fit, then evaluation on a sqrt(m) x sqrt(m) grid."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import polatory_b200 as pb
from polatory_b200.operator import Model, Operator, solve
from polatory_b200.ras import RasPreconditioner
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m_side = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-4
rng = np.random.default_rng(0)
pts = rng.uniform(0, 1, (n, 2))
vals = np.sin(2 * np.pi * pts[:, 0]) * np.cos(3 * np.pi * pts[:, 1]) + 0.5 * np.sin(7 * pts[:, 0] + 5 * pts[:, 1])
rbf = pb.make_rbf("bh2", [1.0, 0.0], 2)
model = Model(rbf, poly_degree=1, nugget=0.0)
bbox = pb.Bbox(np.zeros(2), np.ones(2))
torch.cuda.synchronize(); t0 = time.time()
acc = float(sys.argv[4]) if len(sys.argv) > 4 else tol / 100.0
op = Operator(model, bbox, accuracy=acc); op.set_points(pts)
t1 = time.time()
pc = RasPreconditioner(model, pts, verbose=True)
torch.cuda.synchronize(); t2 = time.time()
w, iters = solve(op, vals, tol, 100, preconditioner=pc.apply)
torch.cuda.synchronize(); t3 = time.time()
print(f"C5 fit: n={n} operator {t1-t0:.2f}s RAS setup {t2-t1:.2f}s solve {t3-t2:.2f}s ({iters} it, {pc.n_levels} levels) "
      f"total {t3-t0:.2f}s config {op.a[0].config()} torch mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
g = np.linspace(0, 1, m_side)
grid = np.ascontiguousarray(np.stack(np.meshgrid(g, g, indexing="ij"), axis=-1).reshape(-1, 2))
ev = pb.make_fmm_evaluator(rbf, bbox)
wh = w.cpu().numpy()
t4 = time.time()
ev.set_source_points(pts); ev.set_weights(wh[:n]); ev.set_target_points(grid)
pred = ev.evaluate() + wh[n] + grid @ wh[n + 1:]
t5 = time.time()
print(f"evaluate {len(grid)} grid targets: {t5-t4:.3f}s  {ev.config()} {{k: round(v, 2) for k, v in ev.phase_times().items()}}", flush=True)
print({k: round(v, 2) for k, v in ev.phase_times().items()})
from oracle import fmm as ofmm
sub = rng.choice(n, 200, replace=False)
fit = ofmm.direct("bh2", [1.0, 0.0], 2, 0, pts, pts[sub], wh[:n]) + wh[n] + pts[sub] @ wh[n + 1:]
print("fit residual on 200 exact samples:", np.max(np.abs(fit - vals[sub])))
