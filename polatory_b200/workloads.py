"""Deterministic synthetic inputs for the BASELINE.json configurations.

The generators restate the *shape* of the reference's own generators --
src/point_cloud/random_points.cpp:7-59 (uniform points on a sphere surface / in a cuboid),
src/point_cloud/sdf_data_generator.cpp:43-99 (surface points + two normal-offset points per
surface point with values 0, +d, -d) -- with numpy's PCG64 instead of std::mt19937, so they
are reproducible here but not bit-identical to a C++ run of the reference.
"""
from __future__ import annotations

import numpy as np


def sphere_surface_points(n, seed=0, radius=1.0):
    """random_points(Sphere3, n, seed): uniform on the sphere surface."""
    rng = np.random.default_rng(seed)
    v = rng.standard_normal((n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return radius * v


def cuboid_points(n, dim=3, seed=0, lo=0.0, hi=1.0):
    """random_points(Cuboid3, n, seed): uniform in a box."""
    rng = np.random.default_rng(seed)
    return rng.uniform(lo, hi, (n, dim))


def sdf_offset_cloud(n_total, seed=0, offset=1e-2):
    """C2/C3 sources: n_total/3 unit-sphere surface points with exact normals, plus the
    +-offset points along the normals (values 0, +offset, -offset)."""
    n_surf = (n_total + 2) // 3
    p = sphere_surface_points(n_surf, seed)
    normals = p.copy()
    pts = np.concatenate([p, p + offset * normals, p - offset * normals])
    vals = np.concatenate([np.zeros(n_surf), np.full(n_surf, offset), np.full(n_surf, -offset)])
    return np.ascontiguousarray(pts), vals


def uniform_weights(n, seed=1):
    """VecX::Random-like weights, uniform in [-1, 1]."""
    return np.random.default_rng(seed).uniform(-1.0, 1.0, n)


def grid_points(bbox_min, bbox_max, shape):
    """Regular grid over a box, row-major, as the isosurface lattice / evaluation grids."""
    axes = [np.linspace(bbox_min[a], bbox_max[a], shape[a]) for a in range(len(shape))]
    mesh = np.meshgrid(*axes, indexing="ij")
    return np.ascontiguousarray(np.stack([m.reshape(-1) for m in mesh], axis=1))


def c3_isosurface_field(n_sources=1_000_000, grid=(216, 216, 215), seed=0):
    """Config #3: 1M-source biharmonic3d interpolant sampled at ~10M grid points over
    1.1 x bbox.  Returns (sources, weights, targets, bbox_min, bbox_max)."""
    src, _ = sdf_offset_cloud(n_sources, seed)
    w = uniform_weights(len(src), seed + 1)
    lo, hi = src.min(axis=0), src.max(axis=0)
    c, half = 0.5 * (lo + hi), 0.5 * (hi - lo)
    tlo, thi = c - 1.1 * half, c + 1.1 * half
    trg = grid_points(tlo, thi, grid)
    return src, w, trg, tlo, thi
