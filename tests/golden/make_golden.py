"""Generates tests/golden/direct_golden.npz.

The reference holds no golden vectors for the FMM path and cannot be built or imported here
(C++, un-vendored dependencies), so these vectors are produced by the numpy
oracle (oracle/direct.py) itself: they freeze the oracle against regressions and give the GPU
tests fixed inputs/outputs that travel to the GPU box.  They are NOT outputs of the reference.
Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import ALL_RBFS, default_params, random_anisotropy  # noqa: E402
from oracle import direct as odir  # noqa: E402
from oracle import rbf as orbf  # noqa: E402


def main():
    rng = np.random.default_rng(20261017)
    out = {}
    cases = []
    for dim in (1, 2, 3):
        for name in ALL_RBFS:
            a = random_anisotropy(dim, rng)
            for kind in range(4):
                if kind == 3 and name in ("sph", "cub"):
                    continue
                src = rng.uniform(-1, 1, (48, dim))
                trg = rng.uniform(-1, 1, (32, dim))
                w = rng.uniform(-1, 1, 48 * odir.kind_km(kind, dim))
                r = orbf.make_rbf(name, default_params(name), dim, a)
                with np.errstate(all="ignore"):
                    ref = odir.full_direct(r, kind, src, trg, w)
                key = f"{name}_{dim}_{kind}"
                cases.append(key)
                out[key + "_aniso"] = a
                out[key + "_src"] = src
                out[key + "_trg"] = trg
                out[key + "_w"] = w
                out[key + "_out"] = ref
    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "direct_golden.npz"), **out)
    print(len(cases), "cases written")


if __name__ == "__main__":
    main()
