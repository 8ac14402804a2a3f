"""ctypes binding of the C ABI declared in include/polatory_b200.h.

The shared library is built in-tree (polatory_b200/libpolatory_b200.so, see
__graft_entry__.build()).  There is no fallback of any kind: if the library is
missing, or no CUDA device is usable, the product path raises.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PLT_B200_LIB: an alternative build of the same library (A/B runs of kernel variants on the GPU box)
LIB_PATH = os.environ.get("PLT_B200_LIB") or os.path.join(_HERE, "libpolatory_b200.so")

PLT_OK, PLT_ERR_INVALID, PLT_ERR_CUDA, PLT_ERR_ACCURACY, PLT_ERR_UNSUPPORTED = 0, 1, 2, 3, 4

# Every symbol include/polatory_b200.h declares: (name, restype, argtypes)
_c_double_p = ctypes.POINTER(ctypes.c_double)
_vp = ctypes.c_void_p


class PltConfig(ctypes.Structure):
    _fields_ = [("tree_height", ctypes.c_int), ("order", ctypes.c_int), ("d", ctypes.c_int)]


SYMBOLS = [
    ("plt_eval_create", ctypes.c_int,
     [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, ctypes.c_int, _vp, _vp, _vp,
      ctypes.POINTER(_vp)]),
    ("plt_eval_destroy", None, [_vp]),
    ("plt_eval_set_source_points", ctypes.c_int, [_vp, _vp, ctypes.c_int64]),
    ("plt_eval_set_target_points", ctypes.c_int, [_vp, _vp, ctypes.c_int64]),
    ("plt_eval_set_points", ctypes.c_int, [_vp, _vp, ctypes.c_int64]),
    ("plt_eval_set_weights", ctypes.c_int, [_vp, _vp, ctypes.c_int64]),
    ("plt_eval_set_accuracy", ctypes.c_int, [_vp, ctypes.c_double]),
    ("plt_eval_evaluate", ctypes.c_int, [_vp, _vp, ctypes.c_int64]),
    ("plt_eval_evaluate_points", ctypes.c_int, [_vp, _vp, ctypes.c_int64, _vp, ctypes.c_int64]),
    ("plt_eval_force_config", ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    ("plt_eval_force_direct", ctypes.c_int, [_vp, ctypes.c_int]),
    ("plt_eval_get_config", ctypes.c_int, [_vp, ctypes.POINTER(PltConfig)]),
    ("plt_eval_set_stream", ctypes.c_int, [_vp, _vp]),
    ("plt_eval_set_target_shard", ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int]),
    ("plt_eval_set_partition", ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp]),
    ("plt_tree_height", ctypes.c_int, [ctypes.c_int, ctypes.c_int64]),
    ("plt_eval_point_keys", ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int, _vp]),
    ("plt_eval_allgather_count", ctypes.c_int64, [_vp]),
    ("plt_eval_get_permutation", ctypes.c_int, [_vp, _vp, ctypes.c_int64]),
    ("plt_eval_get_target_shard_range", ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_int64),
                                                       ctypes.POINTER(ctypes.c_int64)]),
    ("plt_eval_gram_batched", ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_double, _vp]),
    ("plt_eval_gram_mixed", ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_double, _vp]),
    ("plt_eval_phase_times", ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_char_p), _c_double_p, ctypes.c_int]),
    ("plt_eval_work_stats", ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64),
                                           ctypes.POINTER(ctypes.c_int64)]),
    ("plt_eval_launch_count", ctypes.c_int64, [_vp]),
    ("plt_last_error", ctypes.c_char_p, [_vp]),
    ("plt_set_block_m2l_min_fill", ctypes.c_double, [ctypes.c_double]),
    ("plt_set_hadamard_tmem", ctypes.c_int, [ctypes.c_int]),
    ("plt_release_cached_memory", ctypes.c_int64, []),
    ("plt_cached_memory", ctypes.c_int64, []),
    ("plt_set_cached_memory_limit", None, [ctypes.c_int64]),
    ("plt_measure_fp64_peak", ctypes.c_int, [_c_double_p]),
    ("plt_fgmres_create", ctypes.c_int, [ctypes.c_int64, ctypes.c_int, ctypes.POINTER(_vp)]),
    ("plt_fgmres_destroy", None, [_vp]),
    ("plt_fgmres_set_operator", ctypes.c_int, [_vp, _vp, _vp]),
    ("plt_fgmres_set_right_preconditioner", ctypes.c_int, [_vp, _vp, _vp]),
    ("plt_fgmres_set_allreduce", ctypes.c_int, [_vp, _vp, _vp]),
    ("plt_fgmres_set_stream", ctypes.c_int, [_vp, _vp]),
    ("plt_fgmres_setup", ctypes.c_int, [_vp, _vp, _vp]),
    ("plt_fgmres_iterate", ctypes.c_int, [_vp]),
    ("plt_fgmres_solution", ctypes.c_int, [_vp, _vp]),
    ("plt_fgmres_status", ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_int), _c_double_p, _c_double_p]),
    ("plt_fgmres_launch_count", ctypes.c_int64, [_vp]),
    ("plt_fgmres_last_error", ctypes.c_char_p, [_vp]),
    ("plt_ras_choose_coarse_points", ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_int64, _vp, ctypes.c_int64,
                                                    ctypes.c_int64, _vp]),
    ("plt_ras_divide_domains", ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_int64, _vp, ctypes.c_int64,
                                              ctypes.c_int64, ctypes.c_double, ctypes.POINTER(_vp)]),
    ("plt_ras_choose_coarse_points_mixed", ctypes.c_int,
     [_vp, _vp, ctypes.c_int, _vp, ctypes.c_int64, _vp, ctypes.c_int64, _vp, ctypes.c_int64, ctypes.c_int64,
      _vp, ctypes.POINTER(ctypes.c_int64), _vp, ctypes.POINTER(ctypes.c_int64)]),
    ("plt_ras_divide_domains_mixed", ctypes.c_int,
     [_vp, _vp, ctypes.c_int, _vp, ctypes.c_int64, _vp, ctypes.c_int64, _vp, ctypes.c_int64, ctypes.c_int64,
      ctypes.c_double, ctypes.POINTER(_vp)]),
    ("plt_ras_domains_total_grads", ctypes.c_int64, [_vp]),
    ("plt_ras_domains_get_grads", ctypes.c_int, [_vp, _vp, _vp, _vp]),
    ("plt_ras_domains_count", ctypes.c_int64, [_vp]),
    ("plt_ras_domains_total", ctypes.c_int64, [_vp]),
    ("plt_ras_domains_get", ctypes.c_int, [_vp, _vp, _vp, _vp]),
    ("plt_ras_domains_destroy", None, [_vp]),
    ("plt_ras_reduce_q", ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int, ctypes.c_int, _vp, _vp]),
    ("plt_chol_batched", ctypes.c_int, [_vp, ctypes.c_int64, ctypes.c_int, _vp, _vp]),
    ("plt_chol_solve_batched", ctypes.c_int, [_vp, ctypes.c_int64, ctypes.c_int, _vp, ctypes.c_int, _vp, _vp, _vp]),
    ("plt_chol_solve_shared", ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int64, _vp, _vp, _vp]),
    ("plt_gemv", ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp, _vp, _vp]),
    ("plt_ras_sweep_create", ctypes.c_int, [ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_vp)]),
    ("plt_ras_sweep_destroy", None, [_vp]),
    ("plt_ras_sweep_set_level_rows", ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_int64, _vp, ctypes.c_int64]),
    ("plt_ras_sweep_set_fine", ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int64, ctypes.c_int, _vp, _vp, _vp, _vp, _vp,
                                              _vp, ctypes.c_int64]),
    ("plt_ras_sweep_set_coarse", ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp]),
    ("plt_ras_sweep_add_transfer", ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]),
    ("plt_ras_sweep_set_poly", ctypes.c_int, [_vp, _vp, _vp, _vp]),
    ("plt_ras_sweep_apply", ctypes.c_int, [_vp, _vp, _vp, _vp]),
    ("plt_ras_sweep_launch_count", ctypes.c_int64, [_vp]),
    ("plt_ras_sweep_last_error", ctypes.c_char_p, [_vp]),
    ("plt_residual_sample_indices", ctypes.c_int, [_vp, ctypes.c_int64, ctypes.c_int, _vp]),
    ("plt_version", ctypes.c_int, []),
    ("plt_device_check", ctypes.c_int, []),
]

# Callback types of the Krylov solver (include/polatory_b200.h: plt_linop_fn, plt_allreduce_fn).
LINOP_FN = ctypes.CFUNCTYPE(ctypes.c_int, _vp, _vp, _vp)
ALLREDUCE_FN = ctypes.CFUNCTYPE(ctypes.c_int, _vp, _vp, ctypes.c_int)
ALLGATHERV_FN = ctypes.CFUNCTYPE(ctypes.c_int, _vp, _vp, ctypes.POINTER(ctypes.c_int64), ctypes.c_int, _vp)

_lib = None


class PolatoryB200Error(RuntimeError):
    """What the reference reports as std::runtime_error / std::invalid_argument."""

    def __init__(self, status, message):
        super().__init__(message)
        self.status = status


def load():
    """Load the CUDA extension; raises loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(handle, status):
    if status != PLT_OK:
        msg = load().plt_last_error(handle)
        raise PolatoryB200Error(status, msg.decode() if msg else f"status {status}")
