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
    """`grad_sync` hook for FusedAdam: mean of the flat gradient buffer over the data-parallel group, ONE blocking
    all-reduce after backward.  Equal shard sizes make the mean of the per-rank loss means the global-batch mean for
    every term that is a mean over sequences (the cross-entropies, the KL terms); the pairwise latent regulariser pairs
    sequences inside a shard only unless `enable_global_latent_reg` is on (see LatentRegGlobalFn)."""

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


# ------------------------------------------------------------------------------------------------
# overlapped, bucketed gradient all-reduce
# ------------------------------------------------------------------------------------------------
ENCODER_PREFIXES = ("gru_r.", "gru_n.")


def decoder_first(named_params):
    """Order of the live parameters in the flat buffers: everything whose gradient is final BEFORE the encoder BPTT
    starts (decoder recurrences, output / init heads, latent heads, mixture tables) first, the two bidirectional
    encoders last -- so that each group is ONE contiguous range of the flat gradient buffer."""
    early = [(n, p) for n, p in named_params if not n.startswith(ENCODER_PREFIXES)]
    late = [(n, p) for n, p in named_params if n.startswith(ENCODER_PREFIXES)]
    return early + late


class OverlappedGradAllReduce:
    """`grad_sync` hook for FusedAdam that hides most of the gradient all-reduce under the encoder BPTT.

    The flat gradient buffer is cut into two buckets (model.flatten_parameters_ orders it `decoder_first`).  Bucket 0
    (decoder-side parameters, ~2/3 of the bytes at H = 1024) is complete when backward reaches the encoder
    recurrences: a post-accumulate hook on its parameters counts them in and issues the bucket's all-reduce
    asynchronously (NCCL runs it on its own stream, on the SMs the persistent BPTT kernel leaves free), so it runs
    WHILE the encoder BPTT -- the longest single phase of the step -- is computed.  Bucket 1 (the encoders) is reduced
    when FusedAdam.step() calls the hook, which also waits for bucket 0.  The collectives are SUMs: the 1/world factor is
    folded into the backward pass (`loss_scale`, applied by _steps.optimise), so no extra pass over the buffer is needed.
    """

    def __init__(self, model, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.loss_scale = 1.0 / self.world
        self.calls = 0
        self.model = model
        flat, grad = model.flatten_parameters_()
        self.grad = grad
        live = model.live_parameters()
        off, self.split = 0, None
        self._early = []
        for n, p in live:
            if n.startswith(ENCODER_PREFIXES) and self.split is None:
                self.split = off
            if self.split is None:
                self._early.append(p)
            off += ((p.numel() + 3) // 4) * 4
        if self.split is None:
            self.split = off
        self._pending = len(self._early)
        self._handle = None
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self._early] if self.world > 1 else []

    def _on_grad(self, p):
        self._pending -= 1
        if self._pending == 0 and self.split > 0:
            self._handle = dist.all_reduce(self.grad[:self.split], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def __call__(self, flat_grad: torch.Tensor):
        if self.world == 1:
            return flat_grad
        assert flat_grad.data_ptr() == self.grad.data_ptr(), "the flat gradient buffer was re-allocated after the hook was built"
        if self._handle is None:                      # some early parameter received no gradient this step: reduce it now
            if self.split > 0:
                dist.all_reduce(flat_grad[:self.split], op=dist.ReduceOp.SUM, group=self.group)
        if self.split < flat_grad.numel():
            dist.all_reduce(flat_grad[self.split:], op=dist.ReduceOp.SUM, group=self.group)
        if self._handle is not None:
            self._handle.wait()
        self._handle, self._pending = None, len(self._early)
        self.calls += 1
        return flat_grad

    def remove(self):
        for h in self._hooks:
            h.remove()
        self._hooks = []


# ------------------------------------------------------------------------------------------------
# exact global-batch latent regulariser (SURVEY 8e option; reference trainer_gmm.py:199-217 at batch B * world)
# ------------------------------------------------------------------------------------------------
_GLOBAL_LATENT_REG = {"group": None, "on": False}


def enable_global_latent_reg(on: bool = True, group=None):
    """Pair every sequence with every sequence of the GLOBAL batch in the latent regulariser (an all-gather of
    2 x B floats per latent and step) instead of pairing inside a shard."""
    _GLOBAL_LATENT_REG["on"], _GLOBAL_LATENT_REG["group"] = bool(on), group


def global_latent_reg_enabled():
    return _GLOBAL_LATENT_REG["on"] and dist.is_initialized() and dist.get_world_size(_GLOBAL_LATENT_REG["group"]) > 1


class LatentRegGlobalFn(torch.autograd.Function):
    """l = mean over ALL pairs (i, j) of the global batch of (tanh(z_i0 - z_j0) - sign(a_i - a_j))^2.

    Every rank evaluates the same scalar on the gathered latent-dim-0 column and attribute vector (the pairwise kernel
    fn_latent_reg_fwd with Z = 1) and back-propagates only into its own rows.  The data-parallel step averages
    gradients over ranks (1/world), while the true gradient of a term that every rank holds in full is the SUM of the
    per-rank row contributions -- hence the factor `world` on the local rows."""

    @staticmethod
    def forward(ctx, z, attr):
        from . import ops
        group = _GLOBAL_LATENT_REG["group"]
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        B = z.shape[0]
        z0 = z[:, 0].float().contiguous()
        a = attr.to(z.device, dtype=torch.float64).contiguous()
        z_all = torch.empty(world * B, dtype=torch.float32, device=z.device)
        a_all = torch.empty(world * B, dtype=torch.float64, device=z.device)
        dist.all_gather_into_tensor(z_all, z0, group=group)
        dist.all_gather_into_tensor(a_all, a, group=group)
        loss = torch.empty((), dtype=torch.float32, device=z.device)
        dz0 = torch.empty(world * B, dtype=torch.float32, device=z.device)
        rows = torch.empty(world * B, dtype=torch.float32, device=z.device)
        ops.LIB.call("fn_latent_reg_fwd", ops._p(z_all), 1, ops._p(a_all), world * B, ops._p(loss), ops._p(dz0), ops._p(rows), ops._st(z))
        ctx.save_for_backward(dz0[rank * B:(rank + 1) * B].contiguous())
        ctx.meta = (B, z.shape[1], world)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        from . import ops
        (dz0,) = ctx.saved_tensors
        B, Z, world = ctx.meta
        dl = (dloss.float() * world).contiguous()
        dz = torch.empty((B, Z), dtype=torch.float32, device=dz0.device)
        ops.LIB.call("fn_latent_reg_bwd", ops._p(dz0), ops._p(dl), B, Z, ops._p(dz), ops._st(dz))
        return dz, None
