import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gngf_oracle as O
from collision_handling_in_instantngp_b200 import _lib
DEV = "cuda"
def run(P, IN, OUT, leaky=0):
    rng = np.random.default_rng(P + IN)
    enc = (rng.standard_normal((P, IN)) * 0.5).astype(np.float32)
    ws = [(rng.standard_normal((o, i)) / np.sqrt(i)).astype(np.float32) for i, o in [(IN, 64), (64, 64), (64, OUT)]]
    bs = [(rng.standard_normal(o) * 0.1).astype(np.float32) for o in (64, 64, OUT)]
    drgb = rng.standard_normal((P, OUT)).astype(np.float32)
    t = lambda a: torch.from_numpy(a).to(DEV)
    enc_t, ws_t, bs_t, drgb_t = t(enc), [t(w) for w in ws], [t(b) for b in bs], t(drgb)
    rgb = torch.empty((P, OUT), device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    _lib.call("gngf_mlp3_tc_fwd", enc_t.data_ptr(), P, IN, OUT, leaky, ws_t[0].data_ptr(), bs_t[0].data_ptr(),
              ws_t[1].data_ptr(), bs_t[1].data_ptr(), ws_t[2].data_ptr(), bs_t[2].data_ptr(), rgb.data_ptr(), st)
    acts = O.decoder_forward(enc.astype(np.float64), [w.astype(np.float64) for w in ws], [b.astype(np.float64) for b in bs], bool(leaky))
    denc = torch.empty((P, IN), device=DEV)
    gw = [torch.zeros_like(w) for w in ws_t]; gb = [torch.zeros_like(b) for b in bs_t]
    _lib.call("gngf_mlp3_tc_bwd", enc_t.data_ptr(), rgb.data_ptr(), drgb_t.data_ptr(), P, IN, OUT, leaky,
              ws_t[0].data_ptr(), bs_t[0].data_ptr(), ws_t[1].data_ptr(), bs_t[1].data_ptr(), ws_t[2].data_ptr(),
              denc.data_ptr(), gw[0].data_ptr(), gb[0].data_ptr(), gw[1].data_ptr(), gb[1].data_ptr(), gw[2].data_ptr(),
              gb[2].data_ptr(), st)
    torch.cuda.synchronize()
    dz = drgb.astype(np.float64) * acts[-1] * (1 - acts[-1])
    dws, dbs, dx = O._mlp_backward(acts[:-1], [w.astype(np.float64) for w in ws], dz, leaky=bool(leaky))
    err = np.abs(denc.cpu().numpy() - dx).max(1) / np.abs(dx).max()
    ntile = (P + 127) // 128
    pt = np.array([err[i * 128:(i + 1) * 128].max() for i in range(ntile)])
    bad = np.nonzero(pt > 1e-4)[0]
    print(f"P={P} IN={IN}: fwd {np.abs(rgb.cpu().numpy()-acts[-1]).max():.2e} denc max {err.max():.2e}; bad tiles {len(bad)}/{ntile}: {bad[:40]}")
    if len(bad):
        i = bad[0]; e = err[i*128:(i+1)*128]; print("   rows bad in first bad tile:", np.nonzero(e > 1e-4)[0][:40], "vals", e[e>1e-4][:8])
    for i in range(3):
        print("   dW%d err %.2e db%d err %.2e" % (i, np.abs(gw[i].cpu().numpy()-dws[i]).max()/max(np.abs(dws[i]).max(),1e-30), i, np.abs(gb[i].cpu().numpy()-dbs[i]).max()/max(np.abs(dbs[i]).max(),1e-30)))
for P in (1000, 128 * 148, 128 * 149, 128 * 148 * 2, 57404):
    run(P, 8, 3)
run(40000, 32, 3)
