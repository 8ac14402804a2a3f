"""Host-side mirror of the reference's flexible GMRES over the device-resident solver of the C ABI.

  krylov::GmresBase / Gmres / Fgmres   include/polatory/krylov/gmres_base.hpp:11-91,
                                       src/krylov/gmres_base.cpp:7-85, src/krylov/gmres.cpp:9-50,
                                       src/krylov/fgmres.cpp:8-28

Same method names as the reference (`set_initial_solution`, `set_right_preconditioner`, `setup`,
`iterate_process`, `solution_vector`, `relative_residual`, ...).  Vectors are CUDA torch tensors
(float64); the operator and the right preconditioner are callables `f(x, y)` that read the CUDA
tensor `x` and fill the CUDA tensor `y` (both views of the solver's own HBM buffers) on the
current stream.  With `group` set, vectors are this rank's shard and the Arnoldi dot products are
summed across the group with one all_reduce per reduction (NCCL on GPUs).
"""
from __future__ import annotations

import ctypes

from . import _lib


class _DevView:
    """Zero-copy view of `n` doubles at a raw device pointer (CUDA array interface)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def _view(ptr, n):
    import torch
    if n == 0:
        return torch.empty(0, dtype=torch.float64, device="cuda")
    return torch.as_tensor(_DevView(ptr, n), device="cuda")


class Fgmres:
    """krylov::Fgmres(op, rhs, max_iter); `op` is a callable (x, y) or an object with
    `apply(x, y)` and `size()`."""

    def __init__(self, op, rhs, max_iter, group=None):
        import torch
        self._lib = _lib.load()
        self._torch = torch
        self._rhs = rhs.to(torch.float64).contiguous()
        assert self._rhs.is_cuda, "Fgmres works on CUDA tensors"
        self._n = self._rhs.numel()
        self._max_iter = int(max_iter)
        self._x0 = None
        self._group = group
        self._slots = []     # the callables behind the ctypes callbacks, cleared when the solver goes: a ctypes
                             # function pointer sits in a reference cycle of its own, and through the closure it
                             # would keep the operator / preconditioner (and their device memory) alive until a gc pass
        self._err = [None]   # shared with the callbacks (they must not capture `self`: a cycle would keep the
                             # operator, the preconditioner and their device memory alive until a gc pass)
        h = ctypes.c_void_p()
        st = self._lib.plt_fgmres_create(self._n, self._max_iter, ctypes.byref(h))
        if st != _lib.PLT_OK:
            msg = self._lib.plt_fgmres_last_error(None)
            raise _lib.PolatoryB200Error(st, msg.decode() if msg else f"status {st}")
        self._h = h
        self._op_cb = self._wrap(op)
        self._pc_cb = None
        self._check(self._lib.plt_fgmres_set_operator(self._h, ctypes.cast(self._op_cb, ctypes.c_void_p), None))
        self._ar_cb = None
        if group is not None:
            import torch.distributed as dist

            err = self._err

            def allreduce(_ctx, buf, count):
                try:
                    dist.all_reduce(_view(buf, count), op=dist.ReduceOp.SUM, group=group)
                    return 0
                except Exception as e:  # noqa: BLE001 -- reported through the status code
                    err[0] = e
                    return 1

            self._ar_cb = _lib.ALLREDUCE_FN(allreduce)
            self._check(self._lib.plt_fgmres_set_allreduce(self._h, ctypes.cast(self._ar_cb, ctypes.c_void_p), None))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.plt_fgmres_destroy(h)
            self._h = None
        for slot in getattr(self, "_slots", []):
            slot[0] = None

    def _wrap(self, f):
        slot = [f.apply if hasattr(f, "apply") else f]
        self._slots.append(slot)
        n = self._n
        err = self._err

        def cb(_ctx, x, y):
            try:
                slot[0](_view(x, n), _view(y, n))
                return 0
            except Exception as e:  # noqa: BLE001
                err[0] = e
                return 1

        return _lib.LINOP_FN(cb)

    def _check(self, st):
        if st != _lib.PLT_OK:
            if self._err[0] is not None:
                e, self._err[0] = self._err[0], None
                raise e
            msg = self._lib.plt_fgmres_last_error(self._h)
            raise _lib.PolatoryB200Error(st, msg.decode() if msg else f"status {st}")

    # -- reference interface ---------------------------------------------------------
    def set_initial_solution(self, x0):
        assert x0.numel() == self._n
        self._x0 = x0.to(self._torch.float64).contiguous()

    def set_right_preconditioner(self, pc):
        self._pc_cb = self._wrap(pc)
        self._check(self._lib.plt_fgmres_set_right_preconditioner(
            self._h, ctypes.cast(self._pc_cb, ctypes.c_void_p), None))

    def set_left_preconditioner(self, _pc):
        raise RuntimeError("set_left_preconditioner is not supported")  # fgmres.hpp:15-17

    def setup(self):
        x0 = ctypes.c_void_p(self._x0.data_ptr()) if self._x0 is not None and self._n else None
        self._check(self._lib.plt_fgmres_setup(self._h, ctypes.c_void_p(self._rhs.data_ptr() if self._n else 0), x0))

    def iterate_process(self):
        self._check(self._lib.plt_fgmres_iterate(self._h))

    def solution_vector(self):
        x = self._torch.empty(self._n, dtype=self._torch.float64, device=self._rhs.device)
        self._check(self._lib.plt_fgmres_solution(self._h, ctypes.c_void_p(x.data_ptr() if self._n else 0)))
        return x

    def _status(self):
        it, a, r = ctypes.c_int(), ctypes.c_double(), ctypes.c_double()
        self._check(self._lib.plt_fgmres_status(self._h, ctypes.byref(it), ctypes.byref(a), ctypes.byref(r)))
        return it.value, a.value, r.value

    def iteration_count(self):
        return self._status()[0]

    def absolute_residual(self):
        return self._status()[1]

    def relative_residual(self):
        return self._status()[2]

    def max_iterations(self):
        return self._max_iter

    def launch_count(self):
        return int(self._lib.plt_fgmres_launch_count(self._h))
