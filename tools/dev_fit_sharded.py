"""GPU experiment: the config #2 fit on N ranks (torchrun, NCCL): sharded FMM matvec + all-reduced Krylov dots, RAS
replicated on every rank (operator.ShardedPreconditioner)."""
import os, sys, time, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '.')
import polatory_b200 as pb
from polatory_b200 import workloads as wl
from polatory_b200.operator import Model, Operator, ShardedPreconditioner, solve
from polatory_b200.ras import RasPreconditioner
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
tol = 1e-4
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
dist.init_process_group("nccl")
rank, world = dist.get_rank(), dist.get_world_size()
pts, vals = wl.sdf_offset_cloud(n, 0)
model = Model(pb.make_rbf("bh3", [1.0, 0.0]), poly_degree=0, nugget=0.0)
dist.barrier(); torch.cuda.synchronize(); t0 = time.time()
op = Operator(model, pb.Bbox(pts.min(axis=0), pts.max(axis=0)), accuracy=tol / 100.0, group=dist.group.WORLD)
op.set_points(pts)
pc = RasPreconditioner(model, pts)
torch.cuda.synchronize(); t1 = time.time()
rhs_local = op.scatter(np.concatenate([vals, [0.0]]))[:op.hi - op.lo]
w, iters = solve(op, rhs_local, tol, 100, preconditioner=ShardedPreconditioner(op, pc).apply)
torch.cuda.synchronize(); dist.barrier(); t2 = time.time()
full = op.gather(w)
if rank == 0:
    print(f"world={world} n={len(pts)} setup {t1-t0:.2f}s solve {t2-t1:.2f}s ({iters} it) total {t2-t0:.2f}s", flush=True)
    from oracle import fmm as ofmm
    sub = np.random.default_rng(5).choice(len(pts), 200, replace=False)
    wv = full.cpu().numpy()
    fit = ofmm.direct("bh3", [1.0, 0.0], 3, 0, pts, pts[sub], wv[:len(pts)]) + wv[len(pts)]
    print("max |fit - values| on 200 exact samples:", np.max(np.abs(fit - vals[sub])))
dist.destroy_process_group()
