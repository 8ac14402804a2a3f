"""Phase times of the slab-streamed host evaluation against the one-shot one (config #3)."""
import sys, time, json, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import polatory_b200 as pb
from polatory_b200.workloads import c3_isosurface_field

src, w, trg, lo, hi = c3_isosurface_field()
ev = pb.make_fmm_evaluator(pb.make_rbf("bh3", [1.0, 0.0]), pb.Bbox(lo, hi))
ev.set_source_points(torch.from_numpy(src).cuda())
h_trg = torch.from_numpy(trg).pin_memory(); h_w = torch.from_numpy(w).pin_memory()
h_out = torch.empty(len(trg), dtype=torch.float64).pin_memory()
def step():
    ev.set_weights(h_w.numpy())
    ev.evaluate_points(h_trg.numpy(), h_out.numpy())
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10): step()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / 10 * 1e3
ph = ev.phase_times()
print(json.dumps({"slabs": os.environ.get("PLT_SLABS"), "wall_ms": round(wall, 3), "sum_phases": round(sum(ph.values()), 3),
                  "phases": {k: round(v, 3) for k, v in ph.items()}}))
