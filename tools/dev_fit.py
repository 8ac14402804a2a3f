"""GPU experiment: a complete fit -- device FGMRES + RAS preconditioner over the FMM matvec -- of the
config #2 cloud (bh3 SDF centres, degree 0, tolerance 1e-4 absolute, accuracy = tolerance / 100)."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import polatory_b200 as pb
from polatory_b200 import workloads as wl
from polatory_b200.operator import Model, Operator, solve
from polatory_b200.ras import RasPreconditioner
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-4
pts, vals = wl.sdf_offset_cloud(n, 0)
model = Model(pb.make_rbf("bh3", [1.0, 0.0]), poly_degree=0, nugget=0.0)
torch.cuda.synchronize(); t0 = time.time()
op = Operator(model, pb.Bbox(pts.min(axis=0), pts.max(axis=0)), accuracy=tol / 100.0)
op.set_points(pts)
t1 = time.time()
pc = RasPreconditioner(model, pts, verbose=True)
torch.cuda.synchronize(); t2 = time.time()
w, iters = solve(op, vals, tol, 100, preconditioner=pc.apply)
torch.cuda.synchronize(); t3 = time.time()
print(f"n={len(pts)} levels={pc.n_levels} operator setup {t1-t0:.2f}s, RAS setup {t2-t1:.2f}s, solve {t3-t2:.2f}s, "
      f"{iters} iterations, config {op.a[0].config()}, total {t3-t0:.2f}s", flush=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
v = torch.from_numpy(np.concatenate([vals, [0.0]])).cuda(); out = torch.empty_like(v)
pc.apply(v, out); torch.cuda.synchronize(); e0.record(); pc.apply(v, out); e1.record(); torch.cuda.synchronize()
print(f"one RAS application: {e0.elapsed_time(e1):.1f} ms; memory {torch.cuda.max_memory_allocated()/2**30:.1f} GiB torch", flush=True)
# exact residual on a sample
from oracle import fmm as ofmm
sub = np.random.default_rng(5).choice(len(pts), 300, replace=False)
wv = w.cpu().numpy()
fit = ofmm.direct("bh3", [1.0, 0.0], 3, 0, pts, pts[sub], wv[:len(pts)]) + wv[len(pts)]
print("max |fit - values| on 300 exact samples:", np.max(np.abs(fit - vals[sub])))
