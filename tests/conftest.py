import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def random_anisotropy(dim, rng):
    """test/utility.hpp:13-42 with a seeded generator: scaling 10^(+-0.5) (det 1) x rotation."""
    q, _ = np.linalg.qr(rng.standard_normal((dim, dim)))
    if np.linalg.det(q) < 0:
        q[:, 0] *= -1.0
    s = 10.0 ** (0.5 * rng.uniform(-1.0, 1.0, dim))
    s /= np.prod(s) ** (1.0 / dim)
    return np.diag(s) @ q


ALL_RBFS = ["bh3", "th3", "bh2", "th2", "exp", "gau", "gc3", "gc5", "gc7", "gc9", "sp3", "sp5", "sp7", "sp9",
            "sph", "cub"]


def default_params(name):
    return [1.3, 0.1] if name in ("bh3", "th3", "bh2", "th2") else [1.1, 0.7]


@pytest.fixture
def rng():
    return np.random.default_rng(12345)


def dense_th3_hermite(pts, gpts, aniso):
    """mat_a for th3 (phi = r^3, s = 1, c = 0; polyharmonic_odd.hpp:32-67) with anisotropy, rows [values | dim per
    gradient point]: K = phi, F = -(grad phi) A, H = -A^T (Hess phi) A on the transformed differences."""
    dim = pts.shape[1]
    mu, sigma = len(pts), len(gpts)
    tp, tg = pts @ aniso.T, gpts @ aniso.T
    m = mu + dim * sigma
    a = np.zeros((m, m))
    d = tp[:, None, :] - tp[None, :, :]
    a[:mu, :mu] = np.sqrt((d * d).sum(axis=2)) ** 3
    d = tp[:, None, :] - tg[None, :, :]                      # value row i, gradient column j
    r = np.sqrt((d * d).sum(axis=2))
    f = -(3.0 * r[:, :, None] * d) @ aniso                   # -(grad_iso A)
    a[:mu, mu:] = f.reshape(mu, dim * sigma)
    a[mu:, :mu] = a[:mu, mu:].T
    d = tg[:, None, :] - tg[None, :, :]
    r = np.sqrt((d * d).sum(axis=2))
    with np.errstate(all="ignore"):
        hess = 3.0 * (r[:, :, None, None] * np.eye(dim) + np.where(r[:, :, None, None] > 0,
                      d[:, :, :, None] * d[:, :, None, :] / r[:, :, None, None], 0.0))
    h = -np.einsum("ai,pqab,bj->pqij", aniso, hess, aniso)
    a[mu:, mu:] = h.transpose(0, 2, 1, 3).reshape(dim * sigma, dim * sigma)
    return a
