"""GPU experiment: config C4 fit in the reference's configuration (Fitter: matvec at accuracy 0, separate residual
evaluator, RAS) -- th3, anisotropic, n value points + n gradient points in [-1,1]^3, degree 1.
usage: dev_c4_ref.py n tol nugget [max_iter] [transfer_order]"""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import polatory_b200 as pb
from polatory_b200.operator import Model, Solver, monomial_basis
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
tol = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
nugget = float(sys.argv[3]) if len(sys.argv) > 3 else 0.01
max_iter = int(sys.argv[4]) if len(sys.argv) > 4 else 100
torder = int(sys.argv[5]) if len(sys.argv) > 5 else 0
pts = np.random.default_rng(0).uniform(-1, 1, (n, 3)); gpts = np.random.default_rng(1).uniform(-1, 1, (n, 3))
q, _ = np.linalg.qr(np.random.default_rng(2).standard_normal((3, 3)))
A = np.diag(10.0 ** np.array([0.25, 0.0, -0.25])) @ q
f = lambda x: np.sin(np.pi * (x @ A.T)).sum(axis=1)
gradf = lambda x: (np.pi * np.cos(np.pi * (x @ A.T))) @ A
values = np.concatenate([f(pts), gradf(gpts).reshape(-1)])
model = Model(pb.make_rbf("th3", [1.0, 0.0], 3, A), poly_degree=1, nugget=nugget)
torch.cuda.synchronize(); t0 = time.time()
kw = {"transfer_config": (torder, 8 if torder >= 12 else -1)} if torder else {}
solver = Solver(model, pts, gpts, tol / 100, tol / 100, ras_kwargs=kw)
torch.cuda.synchronize(); t1 = time.time()
try:
    w = solver.solve(values, tol, tol, max_iter, verbose=True)
except RuntimeError as e:
    print("FAILED:", e); sys.exit(0)
torch.cuda.synchronize(); t2 = time.time()
pc = solver.pc
print(f"C4 fit: {n}+{n} points, rows {pc.m_rows}, levels {pc.n_levels}, nugget {nugget}, tol {tol}: set-up {t1-t0:.2f}s "
      f"{pc.setup_seconds} solve {t2-t1:.2f}s ({solver.iterations} it) total {t2-t0:.2f}s; matvec configs "
      f"{[e[0].config() for e in (solver.op.a, solver.op.f, solver.op.ft, solver.op.h)]}; residual configs "
      f"{[e[0].config() for e in (solver.res_op.a, solver.res_op.f, solver.res_op.ft, solver.res_op.h)]}; "
      f"torch mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
from oracle import direct as odir, rbf as orbf
o = orbf.make_rbf("th3", [1.0, 0.0], 3, A)
rng = np.random.default_rng(9)
sp, sg = rng.choice(n, 60, replace=False), rng.choice(n, 30, replace=False)
wv = w.cpu().numpy(); m = 4 * n
fit = odir.direct_evaluator(o, 0.0, pts, gpts, wv[:m], pts[sp], gpts[sg]) + monomial_basis(3, 1, pts[sp], gpts[sg]) @ wv[m:]
fit[:60] += nugget * wv[sp]
ref = np.concatenate([values[sp], values[n:].reshape(n, 3)[sg].reshape(-1)])
print("residual on exact samples (values, gradients):", np.max(np.abs(fit[:60] - ref[:60])), np.max(np.abs(fit[60:] - ref[60:])))
