"""torchrun worker of test_dp_gpu.py::test_peer_allreduce_matches_nccl (one rank per GPU, NCCL):
the one-shot NVLink all-reduce (k10_allreduce.cu) against dist.all_reduce, eagerly and inside a CUDA graph."""
import os
import sys

import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    from collision_handling_in_instantngp_b200 import dp
    comm = dp.peer_allreduce_for()
    assert comm is not None, "symmetric memory unavailable"
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    for rep in range(5):                       # both parities of the staging buffer, several epochs
        for n in (1, 3, 4, 1024, 4099, 56212, dp.PEER_MAX_BYTES // 4):
            x = torch.randn(n + 4, device=dev, generator=g)[:n] if n % 4 else torch.randn(n, device=dev, generator=g)
            x = x.contiguous()
            ref = x.clone()
            dist.all_reduce(ref)
            out = comm(x, scale=0.5)
            torch.cuda.synchronize()
            err = (out - 0.5 * ref).abs().max().item()
            assert err <= 1e-5 * max(1.0, ref.abs().max().item()), (rep, n, err)
            gathered = [torch.empty_like(out) for _ in range(world)]
            dist.all_gather(gathered, out)
            assert all(torch.equal(gathered[0], t) for t in gathered), "ranks disagree bitwise"
    # in place + CUDA graph replay
    buf = torch.zeros(56212, device=dev)
    src = torch.randn(56212, device=dev, generator=g)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        buf.copy_(src)
        dp.all_reduce_sum_(buf, 1.0 / world)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        buf.copy_(src)
        dp.all_reduce_sum_(buf, 1.0 / world)
    ref = src.clone()
    dist.all_reduce(ref)
    ref /= world
    for _ in range(7):
        graph.replay()
    torch.cuda.synchronize()
    assert (buf - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item())
    assert not comm.timed_out()
    # latency: peer kernel vs NCCL on the gradient-sized buffer
    def timeit(fn, iters=200):
        for _ in range(20):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters * 1e3
    for n in (1024, 56212, dp.PEER_MAX_BYTES // 4):
        t = torch.randn(n, device=dev)
        us_peer = timeit(lambda: comm(t, out=t, scale=1.0))
        us_nccl = timeit(lambda: dist.all_reduce(t))
        if rank == 0:
            print(f"allreduce {n * 4} B x {world} ranks: peer {us_peer:.1f} us, nccl {us_nccl:.1f} us", flush=True)
    if rank == 0:
        print("PEER_ALLREDUCE_OK", flush=True)
    dist.barrier()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
