"""Data parallelism over pixel batches: one process per GPU, parameters replicated, each rank runs the hot path
on its shard of the coordinates.  Two exchange steps per training step:

  forward   the loss's divergence terms are a non-linear function of the batch-mean slot distribution
            (utils.py:138-144), so the (L, N) column sums are summed over ranks before the loss
            (SyncBatchNorm-style); every rank then holds the same adjoint and no backward collective is needed
            for it (the adjoint of a sum-all-reduce whose consumers are identical is a multiplication by N);
  backward  one NCCL all-reduce (mean) over a single flat buffer of all parameter gradients.

The reference has no distributed code; this is the B200 data-parallel form of functions.py:183-281.
Works with any torch.distributed backend (NCCL on the GPUs; gloo in the CPU tests of the host logic).

On NVLink-connected GPUs both exchanges go through `PeerAllReduce` when the buffer is small (the published
configuration: 4 KB of column sums, ~200 KB of gradients): a one-shot all-reduce kernel over symmetric peer memory
(k10_allreduce.cu) whose cost is one NVLink round trip instead of NCCL's launch + protocol latency; large buffers
(the 268 MB output-layer gradient at T = 2^19) stay with NCCL, which is bandwidth-optimal.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


PEER_MAX_BYTES = 512 << 10        # buffers up to this size use the one-shot peer kernel (measured: 14 vs 19 us at 225 KB,
                                  # 22 vs 18 us at 1 MB on 2 x B200); larger ones NCCL
_PEER = {}                        # (group name, device) -> PeerAllReduce or None (unavailable)


class PeerAllReduce:
    """One-shot all-reduce over NVLink peer memory (k10_allreduce.cu).  Every rank of `group` must construct it and
    call it collectively, with the same sizes in the same order; calls are asynchronous on the current stream and
    can be captured in a CUDA graph."""

    def __init__(self, max_floats: int, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib
        self._lib = _lib
        group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.cap = (int(max_floats) + 3) & ~3
        dev = torch.device("cuda", torch.cuda.current_device())
        self.stage = symm_mem.empty(2 * self.cap, dtype=torch.float32, device=dev)
        self.stage.zero_()
        self.handle = symm_mem.rendezvous(self.stage, group)
        self.stage_ptrs = int(self.handle.buffer_ptrs_dev)
        self.signal_ptrs = int(self.handle.signal_pad_ptrs_dev)
        # slots (block, source rank) of 4 bytes each in the signal pad
        self.max_blocks = max(1, min(64, int(self.handle.signal_pad_size) // 4 // self.world))
        self.state = torch.zeros(4, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        dist.barrier(group)             # every pad / staging buffer is zeroed before the first signal arrives

    def __call__(self, inp: torch.Tensor, out: torch.Tensor = None, scale: float = 1.0) -> torch.Tensor:
        n = inp.numel()
        if n > self.cap or inp.dtype != torch.float32 or not inp.is_contiguous():
            raise ValueError(f"PeerAllReduce: {n} floats exceed the staging capacity {self.cap} (or bad dtype/layout)")
        if out is None:
            out = torch.empty_like(inp)
        self._lib.call("gngf_peer_allreduce", self.stage_ptrs, self.signal_ptrs, self.rank, self.world, inp.data_ptr(),
                       out.data_ptr(), n, self.cap, self.max_blocks, float(scale), self.state.data_ptr(),
                       torch.cuda.current_stream().cuda_stream)
        return out

    def timed_out(self) -> bool:
        return bool(self.state[2].item())

    def check(self) -> None:
        """Raises when a peer failed to reach one of the exchanges within the timeout (k10_allreduce.cu: the flag is
        sticky and the outputs of the affected and all later calls are NaN).  One device->host read."""
        if self.timed_out():
            from ._lib import GngfError
            raise GngfError("PeerAllReduce: a rank did not reach the collective within the timeout; the reduced buffers "
                            "are poisoned (NaN).  Every rank must reach each exchange within the bound "
                            "(set_peer_timeout_ms) -- restart from the last checkpoint")


def set_peer_timeout_ms(ms: int) -> None:
    """How long a rank waits for its peers inside the one-shot all-reduce kernel before declaring the exchange dead
    (default ~30 s).  Eager data-parallel loops that do long rank-local work between steps (checkpoints, host-side
    diagnostics) should raise it."""
    from . import _lib
    _lib.call("gngf_peer_allreduce_set_timeout_ms", int(ms))


def check_exchanges(group=None) -> None:
    """Raises GngfError if any one-shot exchange of this process has timed out (cheap: 4 bytes read back)."""
    for comm in _PEER.values():
        if comm is not None:
            comm.check()


def peer_allreduce_for(group=None):
    """The process-wide PeerAllReduce of `group` (created collectively on first use), or None when symmetric memory
    is unavailable (not NCCL / no peer access / CPU tensors): callers then use the process group's all_reduce."""
    if not dist.is_initialized() or not torch.cuda.is_available():
        return None
    g = group if group is not None else dist.group.WORLD
    if dist.get_backend(g) != "nccl":
        return None
    key = (id(g), torch.cuda.current_device())
    if key not in _PEER:
        ok = torch.zeros(1, device="cuda")
        try:
            comm = PeerAllReduce(PEER_MAX_BYTES // 4, g)
            ok += 1
        except Exception as e:      # noqa: BLE001 -- any failure on any rank must disable it on every rank
            comm = None
            import warnings
            warnings.warn(f"PeerAllReduce unavailable ({type(e).__name__}: {e}); using NCCL all_reduce")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=g)
        _PEER[key] = comm if float(ok.item()) > 0 else None
    return _PEER[key]


def all_reduce_sum_(t: torch.Tensor, scale: float = 1.0, group=None) -> torch.Tensor:
    """In-place scale * sum over ranks of a float32 CUDA tensor: peer kernel when small, NCCL otherwise."""
    comm = peer_allreduce_for(group) if t.is_cuda else None
    if comm is not None and t.dtype == torch.float32 and t.is_contiguous() and t.numel() * 4 <= PEER_MAX_BYTES \
            and t.data_ptr() % 16 == 0:
        comm(t, out=t, scale=scale)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        if scale != 1.0:
            t.mul_(scale)
    return t


def all_reduce_sum(t: torch.Tensor, scale: float = 1.0, group=None) -> torch.Tensor:
    """Out-of-place scale * sum over ranks (no autograd)."""
    comm = peer_allreduce_for(group) if t.is_cuda else None
    if comm is not None and t.dtype == torch.float32 and t.is_contiguous() and t.numel() * 4 <= PEER_MAX_BYTES \
            and t.data_ptr() % 16 == 0:
        return comm(t, scale=scale)
    return all_reduce_sum_(t.clone(), scale, group)


class _AllReduceSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, group):
        ctx.world = dist.get_world_size(group)
        return all_reduce_sum_(t.clone(), 1.0, group)

    @staticmethod
    def backward(ctx, g):
        # every rank evaluates the same function of the reduced tensor, so sum_r' g_r' == world * g
        return g * ctx.world, None


def all_reduce_colsum(colsum: torch.Tensor, group=None) -> torch.Tensor:
    """Sum of the per-rank column sums, differentiable (see module docstring).  A no-op once
    `enable_gradient_allreduce` has installed the in-forward exchange (the module then returns the summed tensor)."""
    from . import ops
    if not dist.is_initialized() or dist.get_world_size(group) == 1 or ops.COLSUM_REDUCE_HOOK is not None:
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


class NodeSharding:
    """Node parallelism for the HPD (last sentence).  Pixel shards alone leave the dominant cost of a
    large configuration replicated: the HPD runs once per lattice NODE, and the nodes a rank's pixels touch overlap
    heavily with its peers' (configs[2]: every rank evaluates all 173 400 nodes whatever its share of the image;
    configs[3]: 2^22 points touch 30 M of the 67 M nodes, an eighth of them still 6 M).  With this helper installed
    (`enable_gradient_allreduce(shard_nodes=True)`, the default) the streaming HPD path instead

      forward   marks the nodes its own points touch, all-gathers the per-rank bitmaps and ORs them (k11: gngf_bitmap_or)
                -> every rank derives the same ascending node list; rank r evaluates rows [r*chunk, (r+1)*chunk) of it
                (layers, bf16 planes, streaming softmax / top-k) and the per-node selections (K values + K indices) are
                all-gathered -- 32 bytes per node for K = 4;
      backward  each rank gathers ITS share of the adjoint of the selected probabilities per listed node, the column-sum
                adjoint folded in (k11: gngf_gather_node_adjoints), a sum-reduce-scatter hands every row to its owner,
                the owner runs the streaming backward and the layer gradients on its rows, and the (partial) HPD
                parameter gradients join the flat gradient all-reduce that already averages tables and decoder.

    No other exchange is added: the column sums and the parameter gradients move as before.  Works on any backend
    (NCCL on the GPUs, gloo in the tests)."""

    ROW_ALIGN = 128          # rows per rank are a multiple of the tensor-core row tile (the last owner takes the rest)

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def chunk_rows(self, total: int) -> int:
        per = -(-int(total) // self.world)
        return max(self.ROW_ALIGN, -(-per // self.ROW_ALIGN) * self.ROW_ALIGN)

    def bounds(self, total: int):
        """(first row, one past the last row, rows per rank) of this rank's share of `total` listed nodes."""
        chunk = self.chunk_rows(total)
        r0 = min(int(total), self.rank * chunk)
        return r0, min(int(total), r0 + chunk), chunk

    def union_bitmap(self, bitmap: torch.Tensor) -> torch.Tensor:
        """OR over ranks of an int32 bitmap (word count a multiple of 4)."""
        words = bitmap.numel()
        maps = torch.empty(self.world * words, dtype=bitmap.dtype, device=bitmap.device)
        dist.all_gather_into_tensor(maps, bitmap.contiguous().reshape(-1), group=self.group)
        if not bitmap.is_cuda:
            maps = maps.reshape(self.world, words)
            out = maps[0].clone()
            for r in range(1, self.world):
                out |= maps[r]
            return out.reshape(bitmap.shape)
        from . import _lib
        out = torch.empty_like(bitmap)
        _lib.call("gngf_bitmap_or", maps.data_ptr(), self.world, bitmap.numel(), out.data_ptr(),
                  torch.cuda.current_stream().cuda_stream)
        return out

    def all_gather_rows(self, local: torch.Tensor, total: int) -> torch.Tensor:
        """local: (chunk, C) whose first (r1 - r0) rows are this rank's -> (total, C): all ranks' rows in list order."""
        out = torch.empty((self.world * local.shape[0], *local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=self.group)
        return out[:total]

    def reduce_scatter_rows(self, padded: torch.Tensor) -> torch.Tensor:
        """padded: (world * chunk, C), rows past the list zero -> (chunk, C): the sum over ranks of this rank's rows."""
        chunk = padded.shape[0] // self.world
        out = torch.empty((chunk, *padded.shape[1:]), dtype=padded.dtype, device=padded.device)
        dist.reduce_scatter_tensor(out, padded.contiguous(), op=dist.ReduceOp.SUM, group=self.group)
        return out


def enable_gradient_allreduce(group=None, shard_nodes: bool = True) -> None:
    """Installs both data-parallel exchanges inside GNGFPath: (1) the parameter gradients are averaged over ranks at
    the end of the backward -- all of them are views of one flat buffer there, so it is a single in-place all-reduce
    and no flatten / unflatten copies (vs. GradientAllReducer, which works on arbitrary parameter lists); (2) the
    (L, N) column sums are summed over ranks inside the forward, on the side stream that produces them, so that the
    exchange overlaps the decoder; `probs.colsum` is then the global sum and `all_reduce_colsum` a no-op."""
    from . import ops

    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        ops.GRAD_REDUCE_HOOK = None
        ops.COLSUM_REDUCE_HOOK = None
        ops.NODE_SHARD = None
        return
    world = dist.get_world_size(group)
    # (3) node parallelism of the streaming HPD path (NodeSharding): large lattices / tables only, eager steps only
    ops.NODE_SHARD = NodeSharding(group) if shard_nodes else None
    peer_allreduce_for(group)       # collective set-up now, not inside the first (possibly graph-captured) step

    def hook(flat: torch.Tensor) -> None:
        all_reduce_sum_(flat, 1.0 / world, group)

    def colsum_hook(colsum: torch.Tensor) -> int:
        all_reduce_sum_(colsum, 1.0, group)
        return world

    ops.GRAD_REDUCE_HOOK = hook
    ops.COLSUM_REDUCE_HOOK = colsum_hook


def shard_bounds(total: int, rank: int, world: int):
    """Even contiguous shards of a batch of `total` points (the last ranks get one fewer when it does not divide)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)
