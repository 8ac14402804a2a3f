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

DEFAULT_CUT_LEVEL = {1: 10, 2: 6, 3: 4}   # 1024 / 4096 / 4096 cells to partition over the ranks


def partition_keys(keys, world_size, dim, level, weights=None):
    """Contiguous Morton key ranges of (near-)equal total weight, cut on level-`level` cell boundaries:
    key_begin (uint32, world_size + 1) with key_begin[0] = 0 and key_begin[-1] = 2^(dim * level).
    `keys`: level-`level` Morton keys of the points to balance (e.g. `evaluator.point_keys`); `weights`:
    optional per-point cost (default 1)."""
    n_cells = 1 << (dim * level)
    hist = np.bincount(np.asarray(keys, dtype=np.int64), weights=weights, minlength=n_cells).astype(np.float64)
    cum = np.concatenate([[0.0], np.cumsum(hist)])
    total = cum[-1]
    kb = np.zeros(world_size + 1, dtype=np.uint32)
    for r in range(1, world_size):
        kb[r] = max(int(np.searchsorted(cum, total * r / world_size, side="left")), int(kb[r - 1]))
    kb[world_size] = n_cells
    return kb


def make_allgatherv(group, rank, owner=None):
    """The `plt_allgatherv_fn` of the C ABI over torch.distributed: in-place all-gather of uneven segments of a
    device buffer.  NCCL: one all_gather whose output tensors are the segments themselves (no staging); otherwise
    (gloo, or an empty segment) segments are padded to the longest one, exchanged with one all_gather and copied
    into place.  Issued on torch's current stream, which is the stream the library works on."""
    import torch
    import torch.distributed as dist
    from .krylov import _view

    cache = {}

    def cb(_ctx, buf, offsets, world, _stream):
        try:
            off = [int(offsets[i]) for i in range(world + 1)]
            lens = [off[r + 1] - off[r] for r in range(world)]
            mx = max(lens)
            if mx == 0:
                return 0
            t = _view(buf, off[-1])
            if dist.get_backend(group) == "nccl" and min(lens) > 0:
                # uneven segments straight into place: ONE grouped NCCL call (ProcessGroupNCCL coalesces the per-rank
                # broadcasts of an all_gather with unequal sizes), no staging copies
                dist.all_gather([t[off[r]:off[r + 1]] for r in range(world)], t[off[rank]:off[rank + 1]].clone(),
                                group=group)
                return 0
            key = (mx, world, t.device)
            if key not in cache:
                cache.clear()
                cache[key] = (torch.zeros(mx, dtype=torch.float64, device=t.device),
                              torch.empty(world * mx, dtype=torch.float64, device=t.device))
            send, recv = cache[key]
            send[:lens[rank]] = t[off[rank]:off[rank + 1]]
            if dist.get_backend(group) == "nccl":
                dist.all_gather_into_tensor(recv, send, group=group)
                parts = [recv[r * mx:(r + 1) * mx] for r in range(world)]
            else:
                parts = [torch.empty_like(send) for _ in range(world)]
                dist.all_gather(parts, send, group=group)
            for r in range(world):
                if r != rank and lens[r]:
                    t[off[r]:off[r + 1]] = parts[r][:lens[r]]
            return 0
        except Exception as e:  # noqa: BLE001 -- reported through the status code
            if owner is not None:
                owner._comm_error = e
            return 1

    return cb


def allgather_shards(local, sizes, group, out=None):
    """Concatenation of the ranks' shards (rank r holds `sizes[r]` doubles) on every rank: a real all-gather
    (padded to the longest shard), e.g. of the weight vector before a sharded matvec."""
    import torch
    import torch.distributed as dist
    world = len(sizes)
    mx = max(sizes)
    send = torch.zeros(mx, dtype=local.dtype, device=local.device)
    send[:local.numel()] = local
    if dist.get_backend(group) == "nccl":
        recv = torch.empty(world * mx, dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(recv, send, group=group)
        parts = [recv[r * mx:r * mx + sizes[r]] for r in range(world)]
    else:
        tmp = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(tmp, send, group=group)
        parts = [tmp[r][:sizes[r]] for r in range(world)]
    if out is None:
        return torch.cat(parts)
    torch.cat(parts, out=out)
    return out


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
