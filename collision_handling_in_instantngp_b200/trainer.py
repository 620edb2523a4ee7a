"""Graph-captured training step: the B200-native way to drive the hot path.

The reference's `train_step` (functions.py:183-281) issues, per batch, ~760 ATen kernels from Python.  The drop-in
module cuts that to ~25 launches, but at the published batch size (57 404 pixels) the GPU work is ~0.35 ms and an
eager Python step is host-bound (~2 ms).  `GraphedTrainer` captures forward + loss + backward + Adam (and, under
data parallelism, the two NCCL all-reduces) once into a CUDA graph over static input buffers; a training step is
then: copy the batch into the static buffers, replay, read the loss.

    trainer = GraphedTrainer(net, optimizer, points=P, gamma=-2, epsilon=1)   # after net.set_coord_bounds(...)
    loss = trainer.step(x_host_pinned, y_host_pinned)                          # or CUDA tensors

The optimizer must be capturable (``optim.FusedAdam``, or ``torch.optim.Adam(..., capturable=True, fused=True)``); the lattice bounds
must be fixed (``net.set_coord_bounds``) because a captured step cannot read the batch's min/max back to the host.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import dp
from ._lib import GngfError
from .loss import fused_loss_and_grads


class GraphedTrainer:
    def __init__(self, net, optimizer, points: int, gamma: float, epsilon: float, l_mse: float = 1.0,
                 l_js_kl: float = 1.0, channels: int = 3, warmup_steps: int = 3, sample_x=None, sample_y=None):
        if net._coord_bounds is None:
            raise GngfError("GraphedTrainer needs fixed lattice bounds: call net.set_coord_bounds(lo, hi) first")
        self.net, self.opt = net, optimizer
        dev = next(net.parameters()).device
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.loss_args = (float(gamma), float(epsilon), float(l_mse), float(l_js_kl))
        self.rows = 4 * points * self.world
        self.x = torch.zeros((points, 2), dtype=torch.float32, device=dev)
        self.y = torch.zeros((points, channels), dtype=torch.float32, device=dev)
        if sample_x is not None:
            self.x.copy_(sample_x)
            self.y.copy_(sample_y)
        else:       # any in-bounds coordinates will do for the warm-up steps
            lo, hi = net._coord_bounds
            self.x.copy_(torch.rand((points, 2), device=dev) * (torch.tensor(hi, device=dev) - torch.tensor(lo, device=dev))
                         + torch.tensor(lo, device=dev))
        if self.world > 1:
            dp.enable_gradient_allreduce()
        self.loss = None
        self.graph = None
        self._capture(warmup_steps)

    def _eager_step(self):
        self.opt.zero_grad(set_to_none=True)
        rgb, probs, _, _ = self.net(self.x, 1.0)
        local = probs.colsum
        # the (L, N) column sums are summed over ranks before the non-linear divergence terms (dp.py); every rank
        # evaluates the same function of the sum, so the adjoint of the local column sums is world * d_colsum
        colsum = dp.all_reduce_sum(local.detach()) if self.world > 1 else local.detach()
        out, d_rgb, d_colsum = fused_loss_and_grads(rgb, self.y, colsum, self.rows, *self.loss_args)
        if self.world > 1:
            d_colsum = d_colsum * float(self.world)
        # the loss kernel emits its own adjoints: they seed the backward directly
        torch.autograd.backward([rgb, local], [d_rgb, d_colsum])
        self.opt.step()
        return out[0]

    def _capture(self, warmup_steps):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup_steps)):
                self._eager_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        self.opt.zero_grad(set_to_none=True)
        self.net.last_state = None
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            self.loss = self._eager_step()
        self.graph = graph

    def load_batch(self, x, y) -> None:
        """Copies a batch (host-pinned or device tensors) into the static input buffers (asynchronous)."""
        self.x.copy_(x, non_blocking=True)
        self.y.copy_(y, non_blocking=True)

    def replay(self) -> torch.Tensor:
        """One training step on the batch currently in the static buffers; returns the (device) loss tensor."""
        self.graph.replay()
        return self.loss

    def step(self, x, y) -> torch.Tensor:
        self.load_batch(x, y)
        return self.replay()
