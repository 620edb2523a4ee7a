import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from collision_handling_in_instantngp_b200 import ops
dev = "cuda"
for M, N, K in [(16384, 65536, 128), (32768, 2**17, 128), (782, 256, 128)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / 11; b = torch.randn(N, device=dev)
    xp, wp = ops.split_bf16x3(x), ops.split_bf16x3(w)
    for _ in range(3): y = ops.tc_linear_fwd(x, w, b, 0, xp, wp)
    torch.cuda.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): y = ops.tc_linear_fwd(x, w, b, 0, xp, wp)
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 5
    fl = 2.0 * M * N * K
    print(f"tc gemm {M}x{N}x{K}: {ms:.3f} ms  useful {fl/ms/1e9:.1f} TFLOP/s  executed(x6) {6*fl/ms/1e9:.1f} TFLOP/s  C write {M*N*4/ms/1e6:.0f} GB/s")
    if M * N <= 2**31:
        ref = torch.addmm(b, x, w.t())
        print("   max rel diff vs torch fp32:", float((y - ref).abs().max() / ref.abs().max()))
import numpy as np
for M, N, K in [(128,128,128),(300,256,128),(1000,1000,64),(4096,8192,128)]:
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((M, K)) * 2).astype(np.float32); w = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32); b = rng.standard_normal(N).astype(np.float32)
    xt, wt, bt = (torch.from_numpy(a).to(dev) for a in (x, w, b))
    ref = x.astype(np.float64) @ w.astype(np.float64).T + b
    y = ops.tc_linear_fwd(xt, wt, bt, 0).cpu().numpy(); y32 = ops.linear_fwd(xt, wt, bt, 0).cpu().numpy(); yt = torch.addmm(bt, xt, wt.t()).cpu().numpy()
    e = lambda a: np.abs(a - ref).max() / np.abs(ref).max()
    print(M, N, K, "tc err %.3e  sgemm err %.3e  torch err %.3e" % (e(y), e(y32), e(yt)))
