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
