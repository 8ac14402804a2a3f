import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import polatory_b200 as pb
from polatory_b200 import workloads as wl
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
g = int(sys.argv[2]) if len(sys.argv) > 2 else 216
order = int(sys.argv[3]) if len(sys.argv) > 3 else 6
src, w, trg, lo, hi = wl.c3_isosurface_field(ns, (g, g, g - 1))
print("sources", len(src), "targets", len(trg))
dev = torch.device("cuda")
tsrc = torch.from_numpy(src).to(dev); tw = torch.from_numpy(w).to(dev); ttrg = torch.from_numpy(trg).to(dev)
out = torch.empty(len(trg), dtype=torch.float64, device=dev)
ev = pb.make_fmm_evaluator(pb.make_rbf("bh3", [1.0, 0.0]), pb.Bbox(lo, hi))
ev.set_source_points(tsrc)
if order != 6: ev.force_config(order, 8 if order >= 12 else -1)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.time()
    ev.set_weights(tw); ev.set_target_points(ttrg); ev.evaluate(out)
    torch.cuda.synchronize(); t1 = time.time()
    pt = ev.phase_times()
    print(f"iter {it}: {1e3*(t1-t0):.2f} ms  {len(trg)/(t1-t0)/1e6:.1f} Mtargets/s", ev.config(), {k: round(v, 3) for k, v in pt.items()}, "sum", round(sum(pt.values()), 2))
# accuracy check on a sample of targets (direct on GPU via a second small evaluator)
sub = np.random.default_rng(5).choice(len(trg), 2000, replace=False)
ev2 = pb.make_fmm_evaluator(pb.make_rbf("bh3", [1.0, 0.0]), pb.Bbox(lo, hi))
ev2.set_source_points(tsrc); ev2.set_weights(tw); ev2.force_config(0, -1, 0)
# brute force by chunks (n_src*n_trg < 2^20 triggers direct): use 1 target at a time is slow; use oracle on a few
from oracle import rbf as orbf, direct as odir
ref = odir.full_direct(orbf.make_rbf("bh3", [1.0, 0.0], 3), 0, src, trg[sub[:200]], w)
got = out.cpu().numpy()[sub[:200]]
print("max abs err", np.max(np.abs(got - ref)), "rel", np.max(np.abs(got - ref)) / np.max(np.abs(ref)), "max|ref|", np.max(np.abs(ref)))
print(torch.cuda.max_memory_allocated() / 1e9, "GB torch;", torch.cuda.mem_get_info())
