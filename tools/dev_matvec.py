"""GPU experiment: the fit's matvec (symmetric evaluator, accuracy 0 -> order 12 / d 8) on the C2 cloud."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import polatory_b200 as pb
from polatory_b200 import workloads as wl
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
acc = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
pts, _ = wl.sdf_offset_cloud(n, 0)
w = wl.uniform_weights(len(pts), 1)
lo, hi = pts.min(axis=0), pts.max(axis=0)
dev = torch.device("cuda")
tp = torch.from_numpy(pts).to(dev); tw = torch.from_numpy(w).to(dev)
out = torch.empty(len(pts), dtype=torch.float64, device=dev)
ev = pb.make_fmm_symmetric_evaluator(pb.make_rbf("bh3", [1.0, 0.0]), pb.Bbox(lo, hi))
ev.set_points(tp); ev.set_accuracy(acc)
for it in range(5):
    torch.cuda.synchronize(); t0 = time.time()
    ev.set_weights(tw); ev.evaluate(out)
    torch.cuda.synchronize(); t1 = time.time()
    pt = ev.phase_times()
    print(f"iter {it}: {1e3*(t1-t0):.2f} ms", ev.config(), {k: round(v, 3) for k, v in pt.items()}, flush=True)
print(ev.work_stats())
from oracle import fmm as ofmm
sub = np.random.default_rng(5).choice(len(pts), 300, replace=False)
ref = ofmm.direct("bh3", [1.0, 0.0], 3, 0, pts, pts[sub], w)
got = out.cpu().numpy()[sub]
print("vs direct: max abs", np.max(np.abs(got - ref)), "rel", np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
print(torch.cuda.mem_get_info())
