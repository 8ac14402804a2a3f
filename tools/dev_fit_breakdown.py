"""Where the wall-time of the solve phase of the config #2 fit goes (bench.py fit_once with synchronised timers)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import polatory_b200 as pb
from polatory_b200 import workloads as wl
from polatory_b200.operator import Model, Operator, ResidualEvaluator
from polatory_b200.ras import RasPreconditioner
from polatory_b200.krylov import Fgmres

points, _ = wl.sdf_offset_cloud(1_000_000, 0)
n = len(points); third = (n + 2) // 3
values = np.concatenate([np.zeros(third), np.full(third, 1e-2), np.full(n - 2 * third, -1e-2)])
tol = 1e-4
dev = torch.device("cuda", 0)

def T():
    torch.cuda.synchronize(); return time.perf_counter()

import gc
for run in range(3):
    free, tot = torch.cuda.mem_get_info(); print('free GB', round(free / 2**30, 1), 'torch reserved GB', round(torch.cuda.memory_reserved() / 2**30, 1))
    model = Model(pb.make_rbf("bh3", [1.0, 0.0]), poly_degree=0, nugget=0.0)
    bbox = pb.Bbox(points.min(axis=0), points.max(axis=0))
    t = [T()]
    op = Operator(model, bbox, 0.0, 0.0); res_op = Operator(model, bbox, tol / 100, tol / 100)
    op.set_points(points); res_op.set_points(points); t.append(T())
    pc = RasPreconditioner(model, points); t.append(T())
    vals = torch.as_tensor(values, dtype=torch.float64).to(dev)
    rhs = torch.cat([vals, torch.zeros(op.local_size() - vals.numel(), dtype=torch.float64, device=dev)])
    solver = Fgmres(op, rhs, 100, group=None); t.append(T())
    solver.set_right_preconditioner(pc.apply); solver.setup(); t.append(T())
    res_eval = ResidualEvaluator(res_op); res_eval.set_values(vals); t.append(T())
    log = []
    while True:
        a = T(); w = solver.solution_vector(); b = T()
        ok, res, gres, exact = res_eval.converged(w, tol, None); c = T()
        if ok: log.append(("sol %.3f conv %.3f (exact=%s)" % (b - a, c - b, exact))); break
        solver.iterate_process(); d = T()
        log.append("sol %.3f conv %.3f iter %.3f" % (b - a, c - b, d - c))
    t.append(T())
    names = ["operators", "ras_setup", "fgmres_create", "fgmres_setup", "residual_eval_create", "loop"]
    print("run", run, {k: round(t[i + 1] - t[i], 3) for i, k in enumerate(names)}, "total", round(t[-1] - t[0], 3))
    print("   ", log, "ras breakdown", {k: round(v, 3) for k, v in pc.setup_seconds.items()})
    del solver, pc, op, res_op, res_eval
    if os.environ.get('DEV_GC'): gc.collect()
