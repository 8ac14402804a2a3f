"""GPU experiment: order-by-order GPU vs oracle vs direct on the failing 3-D shape (run per env toggles)."""
import sys, numpy as np
sys.path.insert(0, '.')
import polatory_b200 as pb
from oracle import fmm as ofmm
rng = np.random.default_rng(12345)
dim, n = 3, 12000
name, params = sys.argv[1] if len(sys.argv) > 1 else "bh3", [1.0, 0.0]
src = rng.uniform(-1, 1, (n, dim)); trg = rng.uniform(-1, 1, (n // 2, dim)); w = rng.uniform(-1, 1, n)
ev = pb.FmmGenericEvaluator(0, pb.make_rbf(name, params, dim), pb.Bbox(-np.ones(dim), np.ones(dim)))
ev.set_source_points(src); ev.set_target_points(trg); ev.set_weights(w)
ref_d = ofmm.direct(name, params, dim, 0, src, trg[:500], w)
sc = np.max(np.abs(ref_d))
for order, d in ((6, -1), (8, -1), (10, -1), (12, -1), (12, 8), (12, 11), (14, 8), (16, 8)):
    ev.force_config(order, d)
    got = ev.evaluate()
    ref = ofmm.fmm(name, params, dim, 0, -np.ones(dim), np.ones(dim), src, trg, w, order, d, 0)
    print(f"order {order:2d} d {d:2d}: gpu-oracle {np.max(np.abs(got-ref))/np.max(np.abs(ref)):.3e}  "
          f"gpu-direct {np.max(np.abs(got[:500]-ref_d))/sc:.3e}  oracle-direct {np.max(np.abs(ref[:500]-ref_d))/sc:.3e}", flush=True)
