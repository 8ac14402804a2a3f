"""Host-side mirror of the reference's FMM evaluator interface, over the C ABI.

  FmmGenericEvaluatorBase<Dim>          include/polatory/fmm/fmm_evaluator.hpp:17-41
  FmmGenericSymmetricEvaluatorBase<Dim> include/polatory/fmm/fmm_symmetric_evaluator.hpp:16-37
  make_fmm_*evaluator factories         include/polatory/fmm/fmm_evaluator.hpp:92-106,
                                        include/polatory/fmm/fmm_symmetric_evaluator.hpp:80-86

Same method names, argument meaning and error behaviour (exceptions carry the reference's
messages).  Points / weights may be numpy arrays (host) or CUDA torch tensors (device
resident, zero-copy across the ABI); `evaluate()` returns a numpy vector, or fills `out`
when a CUDA tensor is passed.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from .rbf import Rbf

KIND_K, KIND_F, KIND_FT, KIND_H = 0, 1, 2, 3
PART_FULL, PART_DIRECT, PART_FAST = 0, 1, 2
kClassic = -1


class Bbox:
    """geometry::Bbox<Dim> (include/polatory/geometry/bbox3d.hpp): min / max corners."""

    def __init__(self, min_, max_):
        self.min = np.asarray(min_, dtype=np.float64).reshape(-1)
        self.max = np.asarray(max_, dtype=np.float64).reshape(-1)

    @staticmethod
    def from_points(points):
        points = np.asarray(points, dtype=np.float64)
        return Bbox(points.min(axis=0), points.max(axis=0))

    def convex_hull(self, other):
        return Bbox(np.minimum(self.min, other.min), np.maximum(self.max, other.max))


def _is_torch_cuda(x):
    return hasattr(x, "is_cuda") and x.is_cuda


def _as_ptr(x, keep):
    """Pointer + element count of a contiguous float64 numpy array or CUDA tensor."""
    if _is_torch_cuda(x):
        import torch
        if x.dtype != torch.float64 or not x.is_contiguous():
            x = x.to(torch.float64).contiguous()
        keep.append(x)
        return ctypes.c_void_p(x.data_ptr()), x.numel()
    a = np.ascontiguousarray(x, dtype=np.float64)
    keep.append(a)
    return ctypes.c_void_p(a.ctypes.data), a.size


class _EvaluatorBase:
    _symmetric = False

    def __init__(self, kind, rbf: Rbf, bbox: Bbox, part=PART_FULL):
        if not isinstance(rbf, Rbf):
            raise RuntimeError("not implemented")  # make_fmm_evaluator.cpp:68
        self._lib = _lib.load()
        self.kind = kind
        self.dim = rbf.dim
        self.km = rbf.dim if kind in (KIND_F, KIND_H) else 1
        self.kn = rbf.dim if kind in (KIND_FT, KIND_H) else 1
        self._n_src = 0
        self._n_trg = 0
        params = np.asarray(rbf.parameters(), dtype=np.float64)
        aniso = np.ascontiguousarray(rbf.anisotropy(), dtype=np.float64)
        bmin = np.ascontiguousarray(bbox.min, dtype=np.float64)
        bmax = np.ascontiguousarray(bbox.max, dtype=np.float64)
        if bmin.size != self.dim or bmax.size != self.dim:
            raise ValueError("bbox dimension mismatch")
        h = ctypes.c_void_p()
        st = self._lib.plt_eval_create(kind, int(self._symmetric), self.dim, rbf.rbf_id, part,
                                       params.ctypes.data, params.size, aniso.ctypes.data,
                                       bmin.ctypes.data, bmax.ctypes.data, ctypes.byref(h))
        if st != _lib.PLT_OK:
            msg = self._lib.plt_last_error(None)
            raise _lib.PolatoryB200Error(st, msg.decode() if msg else f"status {st}")
        self._h = h

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.plt_eval_destroy(h)
            self._h = None

    # -- reference interface ---------------------------------------------------------
    def set_accuracy(self, accuracy):
        _lib.check(self._h, self._lib.plt_eval_set_accuracy(self._h, float(accuracy)))

    def set_weights(self, weights):
        keep = []
        p, n = _as_ptr(weights, keep)
        _lib.check(self._h, self._lib.plt_eval_set_weights(self._h, p, n))

    def evaluate(self, out=None):
        n = self.kn * (self._n_src if self._symmetric else self._n_trg)
        if out is not None and _is_torch_cuda(out):
            import torch
            assert out.dtype == torch.float64 and out.is_contiguous() and out.numel() == n
            _lib.check(self._h, self._lib.plt_eval_evaluate(self._h, ctypes.c_void_p(out.data_ptr()), n))
            return out
        res = np.empty(n, dtype=np.float64) if out is None else out
        assert res.dtype == np.float64 and res.flags.c_contiguous and res.size == n
        _lib.check(self._h, self._lib.plt_eval_evaluate(self._h, ctypes.c_void_p(res.ctypes.data), n))
        return res

    # -- additions (no reference counterpart) ------------------------------------------
    def force_config(self, order, d=kClassic, tree_height=0):
        _lib.check(self._h, self._lib.plt_eval_force_config(self._h, int(order), int(d), int(tree_height)))

    def force_direct(self, on=True):
        """Exact direct summation whatever the size (the residual sample of the fit)."""
        _lib.check(self._h, self._lib.plt_eval_force_direct(self._h, int(bool(on))))

    def config(self):
        c = _lib.PltConfig()
        _lib.check(self._h, self._lib.plt_eval_get_config(self._h, ctypes.byref(c)))
        return {"tree_height": c.tree_height, "order": c.order, "d": c.d}

    def set_stream(self, cuda_stream_ptr):
        _lib.check(self._h, self._lib.plt_eval_set_stream(self._h, ctypes.c_void_p(cuda_stream_ptr)))

    def set_target_shard(self, rank, world_size):
        _lib.check(self._h, self._lib.plt_eval_set_target_shard(self._h, int(rank), int(world_size)))

    def point_keys(self, points, level):
        """Morton keys at `level` of host points (the evaluator's anisotropy and root box applied)."""
        pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, self.dim)
        keys = np.empty(len(pts), dtype=np.uint32)
        _lib.check(self._h, self._lib.plt_eval_point_keys(self._h, ctypes.c_void_p(pts.ctypes.data), len(pts),
                                                          int(level), ctypes.c_void_p(keys.ctypes.data)))
        return keys

    def set_partition(self, rank, world_size, level_cut=0, key_begin=None, group=None, allgatherv=None):
        """Multi-GPU evaluation (include/polatory_b200.h: plt_eval_set_partition): rank `rank` owns the level-
        `level_cut` Morton keys [key_begin[rank], key_begin[rank + 1]); the level-cut multipole expansions are
        exchanged with one all-gather over `group` (torch.distributed; NCCL on GPUs).  world_size <= 1 removes
        the partition."""
        if world_size <= 1:
            _lib.check(self._h, self._lib.plt_eval_set_partition(self._h, 0, 1, 0, None, None, None))
            self._ag_cb = None
            return
        from .parallel import make_allgatherv
        kb = np.ascontiguousarray(key_begin, dtype=np.uint32)
        assert kb.size == world_size + 1
        # `allgatherv`: a replacement callable (ctx, buf, offsets, world, stream) -> status, for tests
        self._ag_cb = _lib.ALLGATHERV_FN(allgatherv if allgatherv is not None else make_allgatherv(group, rank, self))
        _lib.check(self._h, self._lib.plt_eval_set_partition(
            self._h, int(rank), int(world_size), int(level_cut), ctypes.c_void_p(kb.ctypes.data),
            ctypes.cast(self._ag_cb, ctypes.c_void_p), None))

    def allgather_count(self):
        return int(self._lib.plt_eval_allgather_count(self._h))

    def permutation(self):
        """perm[i] = caller index of the i-th target point in Morton order (builds the tree)."""
        n = self._n_src if self._symmetric else self._n_trg
        perm = np.empty(n, dtype=np.int32)
        _lib.check(self._h, self._lib.plt_eval_get_permutation(self._h, ctypes.c_void_p(perm.ctypes.data), n))
        return perm

    def target_shard_range(self):
        """Sorted-order point range [begin, end) of the current target shard."""
        b, e = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self._h, self._lib.plt_eval_get_target_shard_range(self._h, ctypes.byref(b), ctypes.byref(e)))
        return b.value, e.value

    def gram_batched(self, points, counts, nugget, out):
        """Batched Gram matrices (RAS local problems): CUDA tensors points (B, m, dim) float64,
        counts (B,) int32, out (B, m, m) float64."""
        b, m = int(points.shape[0]), int(points.shape[1])
        assert points.is_cuda and counts.is_cuda and out.is_cuda and points.is_contiguous() and out.is_contiguous()
        assert tuple(out.shape) == (b, m, m) and int(points.shape[2]) == self.dim and counts.numel() == b
        for b0 in range(0, b, 65535):
            b1 = min(b, b0 + 65535)
            _lib.check(self._h, self._lib.plt_eval_gram_batched(
                self._h, ctypes.c_void_p(points[b0:b1].data_ptr()), ctypes.c_void_p(counts[b0:b1].data_ptr()),
                b1 - b0, m, float(nugget), ctypes.c_void_p(out[b0:b1].data_ptr())))
        return out

    def gram_mixed(self, points, types, nugget, out):
        """Batched mat_a with value and gradient rows: CUDA tensors points (B, m, dim) float64 (coordinates of each
        row's point), types (B, m) int8 (0 value, 1 + c gradient component, < 0 padding), out (B, m, m)."""
        b, m = int(points.shape[0]), int(points.shape[1])
        assert points.is_cuda and types.is_cuda and out.is_cuda and points.is_contiguous() and out.is_contiguous()
        assert types.is_contiguous() and tuple(types.shape) == (b, m) and tuple(out.shape) == (b, m, m)
        for b0 in range(0, b, 65535):
            b1 = min(b, b0 + 65535)
            _lib.check(self._h, self._lib.plt_eval_gram_mixed(
                self._h, ctypes.c_void_p(points[b0:b1].data_ptr()), ctypes.c_void_p(types[b0:b1].data_ptr()),
                b1 - b0, m, float(nugget), ctypes.c_void_p(out[b0:b1].data_ptr())))
        return out

    def phase_times(self):
        cap = 32
        names = (ctypes.c_char_p * cap)()
        ms = (ctypes.c_double * cap)()
        n = self._lib.plt_eval_phase_times(self._h, names, ms, cap)
        return {names[i].decode(): ms[i] for i in range(n)}

    def work_stats(self):
        a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self._h, self._lib.plt_eval_work_stats(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"m2l_pairs": a.value, "m2l_target_cells": b.value, "p2p_pairs": c.value}

    def launch_count(self):
        return int(self._lib.plt_eval_launch_count(self._h))

    def _points(self, points):
        keep = []
        p, n = _as_ptr(points, keep)
        if n % self.dim != 0:
            raise ValueError("points must be N x dim")
        return p, n // self.dim, keep


class FmmGenericEvaluator(_EvaluatorBase):
    """FmmGenericEvaluator<Kernel> (src/fmm/fmm_evaluator.hpp:32-293)."""

    def set_source_points(self, points):
        p, n, keep = self._points(points)
        _lib.check(self._h, self._lib.plt_eval_set_source_points(self._h, p, n))
        self._n_src = n

    def set_target_points(self, points):
        p, n, keep = self._points(points)
        _lib.check(self._h, self._lib.plt_eval_set_target_points(self._h, p, n))
        self._n_trg = n

    def evaluate_points(self, points, out=None):
        """set_target_points(points) + evaluate(out) as one call (interpolation/evaluator.hpp:83-87): host arrays
        are streamed through the device in slabs (plt_eval_evaluate_points); same values, bit for bit."""
        p, n, keep = self._points(points)
        m = self.kn * n
        if out is not None and _is_torch_cuda(out):
            import torch
            assert out.dtype == torch.float64 and out.is_contiguous() and out.numel() == m
            q = ctypes.c_void_p(out.data_ptr())
            res = out
        else:
            res = np.empty(m, dtype=np.float64) if out is None else out
            assert res.dtype == np.float64 and res.flags.c_contiguous and res.size == m
            q = ctypes.c_void_p(res.ctypes.data)
        _lib.check(self._h, self._lib.plt_eval_evaluate_points(self._h, p, n, q, m))
        self._n_trg = n
        return res


class FmmGenericSymmetricEvaluator(_EvaluatorBase):
    """FmmGenericSymmetricEvaluator<Kernel> (src/fmm/fmm_symmetric_evaluator.hpp:31-275)."""

    _symmetric = True

    def set_points(self, points):
        p, n, keep = self._points(points)
        _lib.check(self._h, self._lib.plt_eval_set_points(self._h, p, n))
        self._n_src = n


# -- the six factories ------------------------------------------------------------------
def tree_height(dim, n_points):
    """src/fmm/utility.hpp:12-16."""
    return int(_lib.load().plt_tree_height(int(dim), int(n_points)))


def make_fmm_evaluator(rbf, bbox):
    return FmmGenericEvaluator(KIND_K, rbf, bbox)


def make_fmm_gradient_evaluator(rbf, bbox):
    return FmmGenericEvaluator(KIND_F, rbf, bbox)


def make_fmm_gradient_transpose_evaluator(rbf, bbox):
    return FmmGenericEvaluator(KIND_FT, rbf, bbox)


def make_fmm_hessian_evaluator(rbf, bbox):
    return FmmGenericEvaluator(KIND_H, rbf, bbox)


def make_fmm_symmetric_evaluator(rbf, bbox):
    return FmmGenericSymmetricEvaluator(KIND_K, rbf, bbox)


def make_fmm_hessian_symmetric_evaluator(rbf, bbox):
    return FmmGenericSymmetricEvaluator(KIND_H, rbf, bbox)
