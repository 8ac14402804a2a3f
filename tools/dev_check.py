import sys, time, numpy as np
sys.path.insert(0, '.')
import polatory_b200 as pb
from oracle import rbf as orbf, direct as odir
rng = np.random.default_rng(0)
def rand_aniso(dim):
    q, _ = np.linalg.qr(rng.standard_normal((dim, dim)))
    if np.linalg.det(q) < 0: q[:, 0] *= -1
    s = 10 ** (0.5 * rng.uniform(-1, 1, dim)); s /= np.prod(s) ** (1 / dim)
    return np.diag(s) @ q
worst = 0
for dim in (1, 2, 3):
    for name in orbf.RBF_NAMES:
        params = [1.3, 0.1] if name in ("bh3","th3","bh2","th2") else [1.1, 0.7]
        A = rand_aniso(dim)
        for kind in range(4):
            if kind == 3 and name in ("sph", "cub"): continue
            ns, nt = 300, 200
            src = rng.uniform(-1, 1, (ns, dim)); trg = rng.uniform(-1, 1, (nt, dim))
            km = odir.kind_km(kind, dim)
            w = rng.uniform(-1, 1, ns * km)
            o = orbf.make_rbf(name, params, dim, A)
            ref = odir.full_direct(o, kind, src, trg, w)
            r = pb.make_rbf(name, params, dim, A)
            ev = pb.FmmGenericEvaluator(kind, r, pb.Bbox(-np.ones(dim), np.ones(dim)))
            ev.set_source_points(src); ev.set_target_points(trg); ev.set_weights(w)
            got = ev.evaluate()
            err = np.max(np.abs(got - ref)) / max(1e-300, np.max(np.abs(ref)))
            worst = max(worst, err)
            if err > 1e-12: print("MISMATCH", dim, name, kind, err)
print("brute-force worst rel err", worst)
# FMM vs direct
for dim, n in ((3, 20000), (2, 20000), (1, 5000)):
  for name, params in (("bh3", [1.0, 0.0]), ("exp", [1.0, 0.5]), ("th3", [1.0, 0.01])):
    for kind in range(4):
        A = rand_aniso(dim)
        src = rng.uniform(-1, 1, (n, dim)); trg = rng.uniform(-1, 1, (n // 2, dim))
        km = odir.kind_km(kind, dim)
        w = rng.uniform(-1, 1, n * km)
        r = pb.make_rbf(name, params, dim, A)
        ev = pb.FmmGenericEvaluator(kind, r, pb.Bbox(-np.ones(dim), np.ones(dim)))
        ev.set_source_points(src); ev.set_target_points(trg); ev.set_weights(w)
        res = {}
        for order, d in ((6, -1), (10, -1), (12, 8)):
            ev.force_config(order, d)
            t0 = time.time(); got = ev.evaluate(); t1 = time.time()
            sub = rng.choice(n // 2, 300, replace=False)
            o = orbf.make_rbf(name, params, dim, A)
            ref = odir.full_direct(o, kind, src, trg[sub], w)
            kn = odir.kind_kn(kind, dim)
            g = got.reshape(-1, kn)[sub].reshape(-1)
            res[(order, d)] = np.max(np.abs(g - ref)) / np.max(np.abs(ref))
        print(dim, name, kind, ev.config(), {k: f"{v:.2e}" for k, v in res.items()}, ev.phase_times())
