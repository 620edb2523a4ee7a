"""Data parallelism over pixel batches: one process per GPU, parameters replicated, each rank runs the hot path
on its shard of the coordinates.  Two exchange steps per training step:

  forward   the loss's divergence terms are a non-linear function of the batch-mean slot distribution
            (utils.py:138-144), so the (L, N) column sums are summed over ranks before the loss
            (SyncBatchNorm-style); every rank then holds the same adjoint and no backward collective is needed
            for it (the adjoint of a sum-all-reduce whose consumers are identical is a multiplication by N);
  backward  one NCCL all-reduce (mean) over a single flat buffer of all parameter gradients.

The reference has no distributed code; this is the B200 data-parallel form of functions.py:183-281.
Works with any torch.distributed backend (NCCL on the GPUs; gloo in the CPU tests of the host logic).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class _AllReduceSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, group):
        ctx.world = dist.get_world_size(group)
        out = t.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        # every rank evaluates the same function of the reduced tensor, so sum_r' g_r' == world * g
        return g * ctx.world, None


def all_reduce_colsum(colsum: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of the per-rank column sums, differentiable (see module docstring)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return colsum
    return _AllReduceSum.apply(colsum, group)


class GradientAllReducer:
    """Averages the gradients of `params` over ranks through one flat buffer."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view(p.shape))
            off += p.numel()

    def nbytes(self) -> int:
        return self.flat.numel() * 4

    @torch.no_grad()
    def __call__(self) -> None:
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return
        world = dist.get_world_size(self.group)
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        self.flat.div_(world)
        for v, p in zip(self.views, self.params):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)


def enable_gradient_allreduce(group=None) -> None:
    """Averages the parameter gradients over ranks inside GNGFPath.backward: all of them are views of one flat
    buffer there, so it is a single in-place all-reduce and no flatten / unflatten copies (vs. GradientAllReducer,
    which works on arbitrary parameter lists)."""
    from . import ops

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        ops.GRAD_REDUCE_HOOK = None
        return
    world = dist.get_world_size(group)

    def hook(flat: torch.Tensor) -> None:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)

    ops.GRAD_REDUCE_HOOK = hook


def shard_bounds(total: int, rank: int, world: int):
    """Even contiguous shards of a batch of `total` points (the last ranks get one fewer when it does not divide)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
