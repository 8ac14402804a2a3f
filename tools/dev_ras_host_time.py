import sys, time, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from polatory_b200 import workloads as wl
from polatory_b200.ras import choose_coarse_points, divide_domains, level_structure
pts, _ = wl.sdf_offset_cloud(1_000_000, 0)
n = len(pts)
nl, counts = level_structure(n)
print(nl, counts)
idcs = np.arange(n, dtype=np.int64)
poly = [0]
point_idcs = {nl - 1: idcs}
for level in range(nl - 1, 0, -1):
    t0 = time.perf_counter()
    off, rows, inner = divide_domains(pts, point_idcs[level], poly, flat=True)
    t1 = time.perf_counter()
    point_idcs[level - 1] = choose_coarse_points(pts, point_idcs[level], poly, counts[level - 1])
    t2 = time.perf_counter()
    print(f"level {level}: n={len(point_idcs[level])} domains={len(off)-1} divide {t1-t0:.3f}s coarse {t2-t1:.3f}s")
