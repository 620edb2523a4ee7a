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
from . import ops
from .loss import fused_loss_and_grads_split


class GraphedTrainer:
    """Whole training step (forward + loss + backward + optimizer, and under data parallelism the two exchanges) captured
    in CUDA graphs over static input buffers.

    Two buffer sets / two graphs alternate so that the host-to-device copy of batch t+1 (on a copy stream) overlaps the
    replay of step t: `step_pipelined` never blocks the host, returns the loss of the PREVIOUS step as a host float
    (None on the first call) and `flush()` returns the last one.  `step` is the blocking form (copy, replay; the
    caller reads the returned device tensor).

    Warm-up: the constructor runs `warmup_steps` real steps (first-use allocations, lazy library loads) before the capture.
    They are side-effect free: parameters, buffers and the optimizer state (moments and step counters) are snapshotted
    before them and restored afterwards -- into the same storages, so the captured pointers stay valid -- which keeps
    the first user step identical to the reference loop's first step and resumed runs' bias corrections intact.

    Errors raised on the device -- a coordinate outside the bounds promised to `net.set_coord_bounds`, a data-parallel
    peer that never reached an exchange -- are read back next to the loss and raise GngfError from
    `step_pipelined` / `flush` (flags every `check_every` steps, and at once when a loss comes back NaN); `step` / `replay`
    users call `check_errors()` at their own cadence."""

    def __init__(self, net, optimizer, points: int, gamma: float, epsilon: float, l_mse: float = 1.0,
                 l_js_kl: float = 1.0, channels: int = 3, warmup_steps: int = 3, sample_x=None, sample_y=None,
                 pipelined: bool = True):
        if net._coord_bounds is None:
            raise GngfError("GraphedTrainer needs fixed lattice bounds: call net.set_coord_bounds(lo, hi) first")
        self.net, self.opt = net, optimizer
        dev = next(net.parameters()).device
        self.dev = dev
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.loss_args = (float(gamma), float(epsilon), float(l_mse), float(l_js_kl))
        self.rows = 4 * points * self.world
        n_sets = 2 if pipelined else 1
        self.xs = [torch.zeros((points, 2), dtype=torch.float32, device=dev) for _ in range(n_sets)]
        self.ys = [torch.zeros((points, channels), dtype=torch.float32, device=dev) for _ in range(n_sets)]
        for x, y in zip(self.xs, self.ys):
            if sample_x is not None:
                x.copy_(sample_x)
                y.copy_(sample_y)
            else:       # any in-bounds coordinates will do for the warm-up steps
                lo, hi = net._coord_bounds
                x.copy_(torch.rand((points, 2), device=dev) * (torch.tensor(hi, device=dev) - torch.tensor(lo, device=dev))
                        + torch.tensor(lo, device=dev))
        if self.world > 1:
            dp.enable_gradient_allreduce()
        self.losses = [None] * n_sets          # device loss tensor of each graph
        self.graphs = [None] * n_sets
        self.cur = 0                           # buffer set of the next step
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.done = [None] * n_sets            # event: the last replay that read buffer set i has finished
        self.loss_host = [torch.zeros(1).pin_memory() for _ in range(n_sets)]
        self.status_host = [torch.zeros(2, dtype=torch.int32).pin_memory() for _ in range(n_sets)]
        self._status_pending = [False] * n_sets
        self.check_every = 32                  # steps between read-backs of the device error flags (step_pipelined)
        self._nsteps = 0
        self.loss_ready = [None] * n_sets      # event: loss_host[i] holds the loss of the last replay of set i
        self._pending = None                   # buffer set whose loss has not been returned yet
        self._capture(warmup_steps)

    # the first buffer set under the names the single-buffer version used
    @property
    def x(self):
        return self.xs[0]

    @property
    def y(self):
        return self.ys[0]

    @property
    def loss(self):
        return self.losses[0]

    @property
    def graph(self):
        return self.graphs[0]

    def _eager_step(self, i: int = 0):
        self.opt.zero_grad(set_to_none=True)
        # the column-sum branch (multiplicities -> column sums -> under data parallelism their all-reduce -> divergence
        # half of the loss) stays on its side stream until the backward needs its adjoint: it overlaps the decoder
        # forward, the MSE half of the loss, the decoder backward and the encoding's point pass
        ops.DEFER_COLSUM_JOIN = True
        ops.WANT_IDX_TOPK = False          # the step never reads the (P,L,4,K) int64 index output: no gather in the graph
        try:
            rgb, probs, _, _ = self.net(self.xs[i], 1.0)
        finally:
            ops.DEFER_COLSUM_JOIN = False
            ops.WANT_IDX_TOPK = True
        state = self.net.last_state
        fork = state.colsum_fork
        # under data parallelism `probs.colsum` already is the sum over ranks: the exchange happens inside the forward,
        # on the side stream that produces the column sums (dp.enable_gradient_allreduce), and the backward scales the
        # adjoint by the world size
        colsum = probs.colsum
        out, d_rgb, d_colsum = fused_loss_and_grads_split(rgb, self.ys[i], colsum, self.rows, *self.loss_args,
                                                          levels_stream=None if fork is None else fork.side)
        # the loss kernel emits its own adjoints: they seed the backward directly (GNGFPath.backward joins the side stream)
        torch.autograd.backward([rgb, colsum], [d_rgb, d_colsum])
        self.opt.step()
        return out[0]

    def _snapshot(self):
        model = {k: v.detach().clone() for k, v in self.net.state_dict().items()}
        opt = {p: {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()}
               for p, st in self.opt.state.items()}
        return model, opt

    @torch.no_grad()
    def _restore(self, snap):
        model, opt = snap
        for k, v in self.net.state_dict().items():
            v.copy_(model[k])
        for p, st in self.opt.state.items():
            old = opt.get(p)
            for k, v in st.items():
                if torch.is_tensor(v):
                    if old is not None and k in old:
                        v.copy_(old[k])
                    else:
                        v.zero_()           # state created by the warm-up: moments and step counter start from zero
                elif old is not None and k in old:
                    st[k] = old[k]
        self.net._err_flag_sticky.zero_()

    def _capture(self, warmup_steps):
        snap = self._snapshot()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup_steps)):
                self._eager_step(0)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._restore(snap)
        torch.cuda.synchronize()
        for i in range(len(self.graphs)):
            if self.world > 1:
                dist.barrier()
            self.opt.zero_grad(set_to_none=True)
            self.net.last_state = None
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                self.losses[i] = self._eager_step(i)
            self.graphs[i] = graph
        torch.cuda.synchronize()

    def load_batch(self, x, y) -> None:
        """Copies a batch (host-pinned or device tensors) into the first static buffer set (asynchronous)."""
        self.xs[0].copy_(x, non_blocking=True)
        self.ys[0].copy_(y, non_blocking=True)

    def replay(self) -> torch.Tensor:
        """One training step on the batch currently in the first buffer set; returns the (device) loss tensor."""
        self.graphs[0].replay()
        return self.losses[0]

    def step(self, x, y) -> torch.Tensor:
        self.load_batch(x, y)
        return self.replay()

    def step_pipelined(self, x, y):
        """Non-blocking step: the batch is copied on the copy stream into the idle buffer set while the previous step
        may still be running, the step is queued behind it, and its loss is copied to pinned host memory.  Returns the
        loss of the previous step (host float; None on the first call)."""
        i = self.cur
        main = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(self.copy_stream):
            if self.done[i] is not None:
                self.copy_stream.wait_event(self.done[i])        # the last replay reading set i is over
            self.xs[i].copy_(x, non_blocking=True)
            self.ys[i].copy_(y, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(self.copy_stream)
        main.wait_event(copied)
        self.graphs[i].replay()
        self.done[i] = torch.cuda.Event()
        self.done[i].record(main)
        self.loss_host[i].copy_(self.losses[i].reshape(1), non_blocking=True)
        if self._nsteps % self.check_every == 0:       # device error flags: 3 tiny copies, amortised over check_every steps
            self.status_host[i].copy_(self._status_dev(), non_blocking=True)
            self._status_pending[i] = True
        self._nsteps += 1
        self.loss_ready[i] = torch.cuda.Event()
        self.loss_ready[i].record(main)
        prev = self._take_pending()
        self._pending = i
        self.cur = (i + 1) % len(self.graphs)
        return prev

    def _take_pending(self):
        if self._pending is None:
            return None
        j = self._pending
        self.loss_ready[j].synchronize()
        self._pending = None
        if self._status_pending[j]:
            self._status_pending[j] = False
            self._raise_on(self.status_host[j])
        loss = float(self.loss_host[j])
        if loss != loss:                # NaN: a poisoned exchange shows up here one step later -- find out why now
            self.check_errors()
        return loss

    def _status_dev(self) -> torch.Tensor:
        """(2,) int32 view: [coordinate out of bounds, peer exchange timed out] -- one persistent device buffer."""
        if getattr(self, "_status", None) is None:
            comm = dp.peer_allreduce_for() if self.world > 1 else None
            # the two flags live in different allocations; a tiny gather kernel-free way: keep both and copy the 4 bytes
            # of each into one staging tensor with two async device-to-device copies inside this call
            self._status = torch.zeros(2, dtype=torch.int32, device=self.dev)
            self._status_src = (self.net._err_flag_sticky, None if comm is None else comm.state[2:3])
        self._status[0:1].copy_(self._status_src[0], non_blocking=True)
        if self._status_src[1] is not None:
            self._status[1:2].copy_(self._status_src[1], non_blocking=True)
        return self._status

    def _raise_on(self, status) -> None:
        oob, dead = int(status[0]), int(status[1])
        if oob:
            raise GngfError("a batch contained coordinates outside the bounds given to net.set_coord_bounds(): they were "
                            "clamped to the lattice box and the step trained on the wrong nodes")
        if dead:
            raise GngfError("a data-parallel peer did not reach an exchange within the timeout: the reduced gradients of "
                            "that step (and every later one) are NaN; restart from the last checkpoint")

    def check_errors(self) -> None:
        """Blocking form of the device-error check for `step` / `replay` users."""
        st = self._status_dev().cpu()
        self._raise_on(st)

    def flush(self):
        """Waits for the last pipelined step; returns its loss (host float) or None."""
        return self._take_pending()
