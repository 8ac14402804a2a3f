"""polatory_b200 -- B200-native RBF fast-multipole evaluator behind Polatory's src/fmm interface.

Only the hot path is here: csrc/ holds the CUDA kernels and the C ABI,
fmm.py / rbf.py mirror the reference's evaluator and RBF interfaces on the host side.
"""
from .fmm import (Bbox, FmmGenericEvaluator, FmmGenericSymmetricEvaluator, KIND_F, KIND_FT, KIND_H, KIND_K,
                  kClassic, make_fmm_evaluator, make_fmm_gradient_evaluator,
                  make_fmm_gradient_transpose_evaluator, make_fmm_hessian_evaluator,
                  make_fmm_hessian_symmetric_evaluator, make_fmm_symmetric_evaluator)
from .rbf import Rbf, make_rbf

# krylov.Fgmres / operator.Operator (the device Krylov driver and the RBF matvec)
# import torch lazily: `from polatory_b200 import krylov, operator`.

__all__ = [
    "Bbox", "FmmGenericEvaluator", "FmmGenericSymmetricEvaluator", "KIND_K", "KIND_F", "KIND_FT", "KIND_H",
    "kClassic", "make_fmm_evaluator", "make_fmm_gradient_evaluator", "make_fmm_gradient_transpose_evaluator",
    "make_fmm_hessian_evaluator", "make_fmm_symmetric_evaluator", "make_fmm_hessian_symmetric_evaluator",
    "Rbf", "make_rbf",
]
