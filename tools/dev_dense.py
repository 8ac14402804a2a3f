"""GPU experiment: evaluation on a DENSE volume cloud (uniform points in a cube) at order 6 / 8, where the M2L tables are
full.  A/B of the parent-block M2L against the list kernel: run with and without PLT_DEBUG_NO_BLK=1."""
import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import polatory_b200 as pb
n_src = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n_trg = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
order = int(sys.argv[3]) if len(sys.argv) > 3 else 6
rng = np.random.default_rng(0)
src = rng.uniform(-1, 1, (n_src, 3)); trg = rng.uniform(-1, 1, (n_trg, 3)); w = rng.uniform(-1, 1, n_src)
dev = torch.device("cuda")
ev = pb.make_fmm_evaluator(pb.make_rbf("bh3", [1.0, 0.0]), pb.Bbox(-np.ones(3), np.ones(3)))
ev.set_source_points(torch.from_numpy(src).to(dev)); ev.set_target_points(torch.from_numpy(trg).to(dev))
ev.force_config(order, -1)
tw = torch.from_numpy(w).to(dev)
out = torch.empty(n_trg, dtype=torch.float64, device=dev)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.time()
    ev.set_weights(tw); ev.evaluate(out)
    torch.cuda.synchronize(); t1 = time.time()
    print(f"iter {it}: {1e3*(t1-t0):.2f} ms", ev.config(), {k: round(v, 3) for k, v in ev.phase_times().items() if v >= 0.1}, flush=True)
print(ev.work_stats())
from oracle import fmm as ofmm
sub = rng.choice(n_trg, 200, replace=False)
ref = ofmm.direct("bh3", [1.0, 0.0], 3, 0, src, trg[sub], w)
got = out.cpu().numpy()[sub]
print("vs direct: rel", np.max(np.abs(got - ref)) / np.max(np.abs(ref)))
