"""GPU experiment: config C5 (without the inequality outer loop) in the reference's fit configuration -- bh2 terrain:
n uniform points in [0,1]^2, z = smooth terrain, degree 1, Solver (matvec at accuracy 0, residual evaluator at
tol/100, RAS), then evaluation on an m_side x m_side grid.
usage: dev_c5_ref.py n m_side tol nugget [max_iter] [accuracy] [transfer_order] [jitter]
jitter > 0: stratified points (one per cell of a sqrt(n) x sqrt(n) grid, jittered by that fraction of the cell) instead
of uniformly random ones."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import polatory_b200 as pb
from polatory_b200.operator import Model, Solver
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
m_side = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-4
nugget = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
max_iter = int(sys.argv[5]) if len(sys.argv) > 5 else 100
acc = float(sys.argv[6]) if len(sys.argv) > 6 else tol / 100
torder = int(sys.argv[7]) if len(sys.argv) > 7 else 0
jitter = float(sys.argv[8]) if len(sys.argv) > 8 else 0.0
rng = np.random.default_rng(0)
if jitter > 0:
    side = int(round(np.sqrt(n))); n = side * side
    gx, gy = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    pts = (np.stack([gx.reshape(-1), gy.reshape(-1)], axis=1) + 0.5 + jitter * rng.uniform(-0.5, 0.5, (n, 2))) / side
    pts = np.ascontiguousarray(pts[rng.permutation(n)])
else:
    pts = rng.uniform(0, 1, (n, 2))
vals = np.sin(2 * np.pi * pts[:, 0]) * np.cos(3 * np.pi * pts[:, 1]) + 0.5 * np.sin(7 * pts[:, 0] + 5 * pts[:, 1])
rbf = pb.make_rbf("bh2", [1.0, 0.0], 2)
model = Model(rbf, poly_degree=1, nugget=nugget)
torch.cuda.synchronize(); t0 = time.time()
kw = {"transfer_config": (torder, 8 if torder >= 12 else -1)} if torder else {}
solver = Solver(model, pts, None, acc, acc, ras_kwargs=kw)
torch.cuda.synchronize(); t1 = time.time()
try:
    w = solver.solve(vals, tol, tol, max_iter, verbose=True)
except Exception as e:
    print("FAILED:", repr(e)); sys.exit(0)
torch.cuda.synchronize(); t2 = time.time()
pc = solver.pc
print(f"C5 fit: n={n} jitter {jitter} transfer order {torder or 6} nugget {nugget} tol {tol} acc {acc}: set-up {t1-t0:.2f}s {pc.setup_seconds} solve {t2-t1:.2f}s "
      f"({solver.iterations} it, {pc.n_levels} levels) total {t2-t0:.2f}s matvec {solver.op.a[0].config()} residual "
      f"{solver.res_op.a[0].config()} torch mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
wh = w.cpu().numpy()
from oracle import fmm as ofmm
sub = rng.choice(n, 200, replace=False)
fit = ofmm.direct("bh2", [1.0, 0.0], 2, 0, pts, pts[sub], wh[:n]) + wh[n] + pts[sub] @ wh[n + 1:] + nugget * wh[sub]
print("fit residual on 200 exact samples:", np.max(np.abs(fit - vals[sub])))
del solver
g = np.linspace(0, 1, m_side)
grid = np.ascontiguousarray(np.stack(np.meshgrid(g, g, indexing="ij"), axis=-1).reshape(-1, 2))
ev = pb.make_fmm_evaluator(rbf, pb.Bbox(np.zeros(2), np.ones(2)))
ev.set_source_points(pts); ev.set_weights(wh[:n])
d_grid = torch.from_numpy(grid).cuda(); out = torch.empty(len(grid), dtype=torch.float64, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t4 = time.time()
    ev.set_target_points(d_grid); ev.evaluate(out)
    torch.cuda.synchronize(); t5 = time.time()
print(f"evaluate {len(grid)} grid targets (device-resident): {t5-t4:.3f}s = {len(grid)/(t5-t4)/1e6:.0f} Mtargets/s "
      f"{ev.config()} { {k: round(v, 2) for k, v in ev.phase_times().items()} }", flush=True)
sub = rng.choice(len(grid), 200, replace=False)
ref = ofmm.direct("bh2", [1.0, 0.0], 2, 0, pts, grid[sub], wh[:n])
got = out.cpu().numpy()[sub]
print("evaluation vs exact sums on 200 grid targets: max abs", np.max(np.abs(got - ref)), "scale", np.max(np.abs(ref)))
