"""Data-parallel plumbing: the path shards over sequences; one process per GPU, the only
collective is ONE all-reduce (sum -> mean) of the flat gradient buffer per step (NCCL over NVLink on
the GPU box; the same code runs on gloo for the CPU tests of the host logic)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world: int):
    """Rows [lo, hi) of the global batch owned by `rank` (equal shards; remainder to the low ranks)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch, rank: int, world: int):
    """Slices every per-sequence item of a dataset tuple (x, r, n, c, r_density, n_density[, ...])."""
    B = len(batch[0])
    lo, hi = shard_bounds(B, rank, world)
    return tuple(t[lo:hi] for t in batch)


class GradAllReduce:
    """`grad_sync` hook for FusedAdam: mean of the flat gradient buffer over the data-parallel group.
    Equal shard sizes make per-rank loss means / world == the global-batch mean (exact DDP semantics)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.calls = 0

    def __call__(self, flat_grad: torch.Tensor):
        if self.world == 1:
            return flat_grad
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        flat_grad.mul_(1.0 / self.world)
        self.calls += 1
        return flat_grad


def broadcast_parameters(model, src: int = 0, group=None):
    """Make every replica start from rank `src`'s weights (flat buffer when on CUDA)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for p in model.state_dict().values():
        dist.broadcast(p, src=src, group=group)
