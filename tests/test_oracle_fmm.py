"""The CPU restatement of the FMM (oracle/fmm_oracle.c) against the exact direct sum: the
reference's own acceptance criterion (test/interpolation/test_evaluator.cpp:70-75,
test_symmetric_evaluator.cpp:58-62)."""
import numpy as np
import pytest

from conftest import random_anisotropy
from oracle import direct as odir
from oracle import fmm as ofmm


@pytest.mark.parametrize("dim,n", [(3, 3000), (2, 3000), (1, 2000)])
@pytest.mark.parametrize("name,params", [("bh3", [1.0, 0.0]), ("th3", [1.0, 0.01]), ("exp", [1.0, 0.5])])
@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_cpu_fmm_converges_to_direct(dim, n, name, params, kind, rng):
    if dim == 1 and kind == 3 and name == "bh3":
        pytest.skip("the 1-D Hessian of |x| vanishes identically (only rounding noise is left)")
    a = random_anisotropy(dim, rng)
    src = rng.uniform(-1, 1, (n, dim))
    trg = rng.uniform(-1, 1, (n // 2, dim))
    w = rng.uniform(-1, 1, n * odir.kind_km(kind, dim))
    ref = ofmm.direct(name, params, dim, kind, src, trg, w, a)
    scale = np.max(np.abs(ref))
    errs = []
    for order, d in ((6, -1), (10, -1)):
        got = ofmm.fmm(name, params, dim, kind, -np.ones(dim), np.ones(dim), src, trg, w, order, d, 0, a)
        errs.append(np.max(np.abs(got - ref)) / scale)
    assert errs[0] < 5e-4
    assert errs[1] < 5e-6


def test_cpu_fmm_symmetric_self_interaction(rng):
    # src/fmm/fmm_symmetric_evaluator.hpp:163-193
    n = 2500
    pts = rng.uniform(-1, 1, (n, 3))
    w = rng.uniform(-1, 1, n)
    ref = ofmm.direct("exp", [1.0, 0.4], 3, 0, pts, None, w, symmetric=True)
    got = ofmm.fmm("exp", [1.0, 0.4], 3, 0, -np.ones(3), np.ones(3), pts, None, w, 10, -1, 0, symmetric=True)
    assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 5e-6


def test_tree_height_rule():
    # src/fmm/utility.hpp:12-16 and the table of SURVEY.md section 8
    assert ofmm.tree_height(3, 10**5) == 6
    assert ofmm.tree_height(3, 10**6) == 7
    assert ofmm.tree_height(3, 10**7) == 8
    assert ofmm.tree_height(2, 5 * 10**6) == 11
    assert ofmm.tree_height(2, 2 * 10**7) == 12
    assert ofmm.tree_height(3, 10) == 2
