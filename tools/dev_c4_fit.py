"""GPU experiment: config C4 fit -- th3, anisotropic, n value points + n gradient points in [-1,1]^3 (Hermite-Birkhoff),
degree 1, FGMRES + RAS."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import polatory_b200 as pb
from polatory_b200.operator import Model, Operator, monomial_basis, solve
from polatory_b200.ras import RasPreconditioner
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-4
acc = float(sys.argv[3]) if len(sys.argv) > 3 else tol / 100
pts = np.random.default_rng(0).uniform(-1, 1, (n, 3)); gpts = np.random.default_rng(1).uniform(-1, 1, (n, 3))
q, _ = np.linalg.qr(np.random.default_rng(2).standard_normal((3, 3)))
A = np.diag(10.0 ** np.array([0.25, 0.0, -0.25])) @ q
f = lambda x: np.sin(np.pi * (x @ A.T)).sum(axis=1)
gradf = lambda x: (np.pi * np.cos(np.pi * (x @ A.T))) @ A
values = np.concatenate([f(pts), gradf(gpts).reshape(-1)])
model = Model(pb.make_rbf("th3", [1.0, 0.0], 3, A), poly_degree=1)
torch.cuda.synchronize(); t0 = time.time()
op = Operator(model, pb.Bbox(-np.ones(3), np.ones(3)), accuracy=acc, grad_accuracy=acc)
op.set_points(pts, gpts)
t1 = time.time()
torder = int(sys.argv[4]) if len(sys.argv) > 4 else 0
pc = RasPreconditioner(model, pts, gpts, verbose=True, transfer_config=(torder, 8 if torder >= 12 else -1) if torder else None)
torch.cuda.synchronize(); t2 = time.time()
w, iters = solve(op, values, tol, 100, preconditioner=pc.apply)
torch.cuda.synchronize(); t3 = time.time()
print(f"C4 fit: {n}+{n} points, rows {pc.m_rows}, levels {pc.n_levels}: operator {t1-t0:.2f}s RAS setup {t2-t1:.2f}s "
      f"{pc.setup_seconds} solve {t3-t2:.2f}s ({iters} it) total {t3-t0:.2f}s; configs "
      f"{[e[0].config() for e in (op.a, op.f, op.ft, op.h)]}; torch mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
from oracle import direct as odir, rbf as orbf
o = orbf.make_rbf("th3", [1.0, 0.0], 3, A)
rng = np.random.default_rng(9)
sp, sg = rng.choice(n, 60, replace=False), rng.choice(n, 30, replace=False)
wv = w.cpu().numpy(); m = 4 * n
fit = odir.direct_evaluator(o, 0.0, pts, gpts, wv[:m], pts[sp], gpts[sg]) + monomial_basis(3, 1, pts[sp], gpts[sg]) @ wv[m:]
ref = np.concatenate([values[sp], values[n:].reshape(n, 3)[sg].reshape(-1)])
print("residual on exact samples (values, gradients):", np.max(np.abs(fit[:60] - ref[:60])), np.max(np.abs(fit[60:] - ref[60:])))
