"""GPU experiment: config C4 matvec -- th3, anisotropic, n value points + n gradient points
(Hermite-Birkhoff): the four evaluator kinds A, F, F^T, H per operator application."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import polatory_b200 as pb
from polatory_b200.operator import Model, Operator
n = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
acc = float(sys.argv[2]) if len(sys.argv) > 2 else float("inf")
rng = np.random.default_rng(0)
pts = rng.uniform(-1, 1, (n, 3)); gpts = np.random.default_rng(1).uniform(-1, 1, (n, 3))
q, _ = np.linalg.qr(np.random.default_rng(2).standard_normal((3, 3)))
A = np.diag(10.0 ** np.array([0.25, 0.0, -0.25])) @ q
model = Model(pb.make_rbf("th3", [1.0, 0.0], 3, A), poly_degree=1)
op = Operator(model, pb.Bbox(-np.ones(3), np.ones(3)), accuracy=acc, grad_accuracy=acc)
op.set_points(pts, gpts)
x = torch.from_numpy(rng.uniform(-1, 1, op.size())).cuda(); y = torch.empty_like(x)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.time(); op.apply(x, y); torch.cuda.synchronize(); t1 = time.time()
    print(f"iter {it}: {1e3*(t1-t0):.2f} ms")
for name, ev in (("A", op.a[0]), ("F", op.f[0]), ("FT", op.ft[0]), ("H", op.h[0])):
    pt = ev.phase_times()
    print(name, ev.config(), round(sum(pt.values()), 2), {k: round(v, 2) for k, v in pt.items()}, ev.work_stats())
# spot check against the exact sums
from oracle import direct as odir, rbf as orbf
o = orbf.make_rbf("th3", [1.0, 0.0], 3, A)
sub = rng.choice(n, 100, replace=False)
xv = x.cpu().numpy()
ref = odir.direct_evaluator(o, 0.0, pts, gpts, xv[:4 * n], pts[sub], gpts[sub])
from polatory_b200.operator import monomial_basis
P = monomial_basis(3, 1, pts[sub], gpts[sub])
ref = ref + P @ xv[4 * n:]
got = y.cpu().numpy()
gi = np.concatenate([sub, (n + 3 * sub[:, None] + np.arange(3)).reshape(-1)])
print("vs direct rel err:", np.max(np.abs(got[gi] - ref)) / np.max(np.abs(ref)))
