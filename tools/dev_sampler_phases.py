"""Per-batch phase times of the isosurface sampler on config #3 (one lattice layer per batch)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import polatory_b200 as pb
from polatory_b200.workloads import c3_isosurface_field
from polatory_b200.evaluator import RbfFieldFunction
from polatory_b200.operator import Model
src, w, trg, lo, hi = c3_isosurface_field()
d_trg = torch.from_numpy(trg).cuda()
field = RbfFieldFunction(Model(pb.make_rbf("bh3", [1.0, 0.0]), poly_degree=-1), src, w)
field.set_evaluation_bbox(pb.Bbox(lo, hi))
per = 216 * 215
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for b in range(216):
        field(d_trg[b * per:(b + 1) * per])
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ph = field.evaluator.phase_times()
    print("rep", rep, "ms/batch", round((t1 - t0) / 216 * 1e3, 4), "phases of the last batch",
          {k: round(v, 4) for k, v in ph.items()}, "sum", round(sum(ph.values()), 4), "launches/batch",
          field.evaluator.launch_count() // (216 * (rep + 1)))
