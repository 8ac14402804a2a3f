"""GPU experiment: per-iteration cost of the device FGMRES over the FMM matvec on the C2 cloud
(1M bh3 centres, degree 0, accuracy 0 -> order 12), 1 rank or N ranks (torchrun, NCCL).
No preconditioner (RAS is SURVEY 8f-2/3, not built): this measures the Krylov driver and the
sharded matvec, not convergence."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, '.')
import polatory_b200 as pb
from polatory_b200 import workloads as wl
from polatory_b200.operator import Model, Operator
from polatory_b200.krylov import Fgmres

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
acc = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
group = None
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group("nccl")
    group = dist.group.WORLD
pts, vals = wl.sdf_offset_cloud(n, 0)
lo, hi = pts.min(axis=0), pts.max(axis=0)
model = Model(pb.make_rbf("bh3", [1.0, 0.0]), poly_degree=0, nugget=0.0)
op = Operator(model, pb.Bbox(lo, hi), accuracy=acc, group=group)
t0 = time.time(); op.set_points(pts); torch.cuda.synchronize(); t1 = time.time()
rhs_g = np.concatenate([vals, [0.0]])
rhs = op.scatter(rhs_g)
s = Fgmres(op, rhs, iters, group=group)
s.setup()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
times = []
for i in range(iters):
    if group is not None: dist.barrier()
    torch.cuda.synchronize(); ev0.record()
    s.iterate_process()
    ev1.record(); torch.cuda.synchronize()
    times.append(ev0.elapsed_time(ev1))
tt = torch.tensor(times, device="cuda")
if group is not None: dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    print(f"world={world} n={len(pts)} local={op.local_size()} set_points {t1-t0:.2f}s config={op.a[0].config()}")
    print("ms/iter:", [round(float(x), 2) for x in tt.tolist()])
    print("rel residual:", s.relative_residual(), "phases", {k: round(v, 3) for k, v in op.a[0].phase_times().items()})
if group is not None: dist.destroy_process_group()
