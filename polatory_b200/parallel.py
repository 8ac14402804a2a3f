"""Multi-GPU host logic: one process per GPU, targets sharded by Morton range.

Bulk evaluation needs no data-path collective: every rank holds all source points and weights
(24 + 8 bytes per source; the upward pass of 10^6 sources is ~1 ms) and evaluates only the
targets of its contiguous Morton range (`plt_eval_set_target_shard`).  The evaluator writes
zeros for targets outside its range, so the full vector is the SUM over ranks -- one
`all_reduce` when the caller wants the assembled result on every rank (NCCL over NVLink on
GPUs; the same code runs on gloo for the CPU tests), or nothing at all when the result stays
sharded (the matvec inside a sharded Krylov solver, the per-layer isosurface sampler).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n, world_size):
    """Point-index boundaries [n*r/W] the C ABI uses for rank r (evaluator.cu shard_leaves);
    the actual cut is moved to the next leaf boundary on the device."""
    return [n * r // world_size for r in range(world_size + 1)]


class ShardedEvaluator:
    """Wraps an evaluator (polatory_b200.fmm) so that `evaluate()` computes this rank's Morton
    shard and, with `assemble=True`, sums the shards across the process group."""

    def __init__(self, evaluator, rank=None, world_size=None, group=None):
        import torch.distributed as dist
        self._dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world_size = dist.get_world_size(group) if world_size is None else world_size
        self.evaluator = evaluator
        evaluator.set_target_shard(self.rank, self.world_size)

    def __getattr__(self, name):
        return getattr(self.evaluator, name)

    def evaluate(self, out=None, assemble=True):
        import torch
        res = self.evaluator.evaluate(out)
        if not assemble or self.world_size == 1:
            return res
        if isinstance(res, np.ndarray):
            t = torch.from_numpy(res)
            self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM, group=self.group)
            return res
        self._dist.all_reduce(res, op=self._dist.ReduceOp.SUM, group=self.group)
        return res
