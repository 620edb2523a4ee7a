"""CPU (gloo, world_size 2): the data-parallel exchange steps of dp.py reproduce the single-process
large-batch gradient -- column sums all-reduced before the non-linear divergence terms, gradients averaged
through one flat buffer."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from collision_handling_in_instantngp_b200 import dp
from collision_handling_in_instantngp_b200.loss import level_divergences

L, N, D = 3, 16, 5


def _toy(theta, A, b, rows_total, reduce_fn):
    """colsum = softplus(A theta) rows (positive, like probability sums); loss = mse-like local term + divergences."""
    colsum = torch.nn.functional.softplus(A @ theta).reshape(L, N)
    local = ((theta * b) ** 2).sum()
    total = reduce_fn(colsum)
    return local + level_divergences(total / rows_total, gamma=-2.0, epsilon=1.0).sum()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7)
    theta0 = torch.randn(D, generator=g)
    As = [torch.randn(L * N, D, generator=g) for _ in range(world)]
    bs = [torch.randn(D, generator=g) for _ in range(world)]
    theta = torch.nn.Parameter(theta0.clone())
    extra = torch.nn.Parameter(torch.ones(3))            # a parameter that receives no gradient on this rank
    loss = _toy(theta, As[rank], bs[rank], rows_total=4.0 * world, reduce_fn=dp.all_reduce_colsum)
    loss.backward()
    red = dp.GradientAllReducer([theta, extra])
    assert red.nbytes() == (D + 3) * 4
    red()
    # the peer (NVLink) all-reduce is a CUDA + NCCL feature: under gloo the helpers fall back to the process group
    assert dp.peer_allreduce_for() is None
    t = torch.full((5,), float(rank + 1))
    s2 = dp.all_reduce_sum(t, scale=0.5)
    assert torch.equal(t, torch.full((5,), float(rank + 1))) and torch.equal(s2, torch.full((5,), 1.5))
    dp.all_reduce_sum_(t, 1.0 / world)
    assert torch.equal(t, torch.full((5,), 1.5))
    out.put((rank, theta.grad.numpy().copy(), extra.grad.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_two_rank_gradients_equal_single_process_large_batch():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single process: mean of the local terms + divergences of the summed column sums
    g = torch.Generator().manual_seed(7)
    theta0 = torch.randn(D, generator=g)
    As = [torch.randn(L * N, D, generator=g) for _ in range(world)]
    bs = [torch.randn(D, generator=g) for _ in range(world)]
    theta = theta0.clone().requires_grad_()
    colsum = sum(torch.nn.functional.softplus(A @ theta).reshape(L, N) for A in As)
    loss = sum(((theta * b) ** 2).sum() for b in bs) / world + \
        level_divergences(colsum / (4.0 * world), gamma=-2.0, epsilon=1.0).sum()
    loss.backward()
    for rank, grad, extra in got:
        np.testing.assert_allclose(grad, theta.grad.numpy(), rtol=2e-5, atol=1e-7)
        np.testing.assert_array_equal(extra, np.zeros(3, dtype=np.float32))


def test_shard_bounds_cover_the_batch():
    for total in (57404, 172212, 2 ** 22, 7):
        for world in (1, 2, 4, 8):
            spans = [dp.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_is_identity():
    t = torch.ones(2, 3, requires_grad=True)
    assert dp.all_reduce_colsum(t) is t


# ---- node parallelism (dp.NodeSharding): host logic on gloo -----------------------------------------------------
def _node_worker(rank, world, port, out):
    """Emulates the node-parallel HPD with torch CPU ops: per-node "HPD" f(u) = (values, indices), evaluated by the
    owner only, all-gathered; per-rank adjoint shares reduce-scattered to the owners."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = dp.NodeSharding()
    U, K = 1000, 4
    g = torch.Generator().manual_seed(11)
    touched = [torch.rand(U, generator=g) < 0.3 for _ in range(world)]        # nodes each rank's points touch
    # bitmap union -> the agreed list
    words = ((U + 31) // 32 + 3) & ~3
    def bitmap_of(mask):
        bits = torch.zeros(words * 32, dtype=torch.int64)
        bits[:U] = mask.long()
        w = (bits.reshape(words, 32) << torch.arange(32)).sum(1)
        return (w & 0xFFFFFFFF).to(torch.int64).where(w < 2 ** 31, w - 2 ** 32).to(torch.int32)
    union = sh.union_bitmap(bitmap_of(touched[rank]))
    want = bitmap_of(torch.stack(touched).any(0))
    assert torch.equal(union, want)
    ids = torch.stack(touched).any(0).nonzero().flatten()
    total = ids.numel()
    r0, r1, chunk = sh.bounds(total)
    assert chunk % dp.NodeSharding.ROW_ALIGN == 0 and chunk * world >= total
    # forward: owner evaluates its rows, everybody gets all rows
    vals_local = torch.zeros((chunk, K))
    vals_local[:r1 - r0] = ids[r0:r1, None].float() * torch.arange(1, K + 1)
    vals_all = sh.all_gather_rows(vals_local, total)
    assert torch.equal(vals_all, ids[:, None].float() * torch.arange(1, K + 1))
    # backward: every rank holds a share of every listed node's adjoint; the owner receives the sum
    shares = [torch.randn(total, K, generator=g) for _ in range(world)]
    padded = torch.zeros((world * chunk, K))
    padded[:total] = shares[rank]
    mine = sh.reduce_scatter_rows(padded)[:r1 - r0]
    torch.testing.assert_close(mine, sum(shares)[r0:r1])
    out.put((rank, r0, r1, total))
    dist.barrier()
    dist.destroy_process_group()


def test_node_sharding_host_logic_on_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_node_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == 0 and got[0][2] == got[1][1] and got[1][2] == got[0][3] == got[1][3]


def test_node_sharding_bounds():
    class _Fake(dp.NodeSharding):
        def __init__(self, rank, world):
            self.rank, self.world, self.group = rank, world, None
    for total in (0, 1, 127, 128, 129, 1156, 173400, 29972809):
        for world in (2, 3, 4, 8):
            spans = [_Fake(r, world).bounds(total) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert len({s[2] for s in spans}) == 1 and spans[0][2] * world >= total
            assert all(s[1] - s[0] <= s[2] for s in spans)
