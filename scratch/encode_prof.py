"""ncu target: the point passes of the encode kernels at a cfg5 sweep point (P = 2^22, L = 16, F = 2, n_max = 2048)."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from collision_handling_in_instantngp_b200 import _lib
from collision_handling_in_instantngp_b200.lattice import build_lattice, level_resolutions
dev = torch.device("cuda")
L, F, P = 16, 2, 2 ** 22
lat = build_lattice(level_resolutions(16, 2048, L))
S = lat.num_level_nodes
g = torch.Generator(device="cuda").manual_seed(65535)
x = torch.rand((P, 2), device=dev, generator=g)
nfeat = torch.randn((S, F), device=dev, generator=g)
enc = torch.empty((P, L * F), device=dev)
denc = torch.randn((P, L * F), device=dev, generator=g)
dnf = torch.zeros((S, F), device=dev)
cnt = torch.zeros(S + 1, dtype=torch.int32, device=dev)
cell = torch.zeros(S, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
for it in range(3):
    _lib.call("gngf_encode_fwd", x.data_ptr(), P, lat, F, nfeat.data_ptr(), enc.data_ptr(), None, None, None, st)
    _lib.call("gngf_encode_fwd", x.data_ptr(), P, lat, F, nfeat.data_ptr(), enc.data_ptr(), cnt.data_ptr(), cell.data_ptr(), cnt[S:].data_ptr(), st)
    _lib.call("gngf_encode_bwd", x.data_ptr(), P, lat, F, denc.data_ptr(), dnf.data_ptr(), st)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("fwd", lambda: _lib.call("gngf_encode_fwd", x.data_ptr(), P, lat, F, nfeat.data_ptr(), enc.data_ptr(), None, None, None, st)),
                 ("fwd+cnt", lambda: _lib.call("gngf_encode_fwd", x.data_ptr(), P, lat, F, nfeat.data_ptr(), enc.data_ptr(), cnt.data_ptr(), cell.data_ptr(), cnt[S:].data_ptr(), st)),
                 ("bwd", lambda: _lib.call("gngf_encode_bwd", x.data_ptr(), P, lat, F, denc.data_ptr(), dnf.data_ptr(), st))):
    a.record()
    for _ in range(10):
        fn()
    b.record(); torch.cuda.synchronize()
    print(name, a.elapsed_time(b) / 10, "ms")
