#!/usr/bin/env python
"""Encoding-only micro-bench sweep (BASELINE.json configs[4]): batch 2^16..2^24 points x table size 2^14..2^22,
L = 16 levels x F = 2 features, GNGF lattice path vs. the plain hash-mod baseline, forward and backward separately.

    python bench_sweep.py [--quick] > profiles/rNN_cfg5_sweep.jsonl

Per (P, T) and mode it times, with CUDA events (median of 20 after 5 warm-ups, inputs far larger than L2 for
P >= 2^20; an L2 flush precedes every timed iteration otherwise):
  gngf  fwd = node pass (K = 4 table rows mixed per level node) + point pass (4 node-feature gathers per level)
        bwd = point pass (4 vector reductions per level) + node pass (table scatter-add + top-k adjoint)
  hash  fwd = encode_hash_fwd (4 table gathers per level at (x ^ y*2654435761) mod T),  bwd = encode_hash_bwd
and reports algorithmic GB/s against the measured HBM copy peak (MEASURED_PEAKS.json).  Algorithmic bytes per
point (hash-mod rows): fwd 8 + L*4*F*4 + L*F*4 = 648 B, bwd 128 + 512 + 8 = 648 B; the GNGF
point passes move the same bytes (node features instead of table rows) and the node passes add
S*(K*(8+F*4)+F*4) B fwd / S*(F*4+K*(8+2*F*4+4)) B bwd.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--n-max", type=int, default=2048)
    args = ap.parse_args()
    import torch

    from collision_handling_in_instantngp_b200 import _lib, ops
    from collision_handling_in_instantngp_b200.lattice import build_lattice, level_resolutions

    dev = torch.device("cuda")
    L, F, K = 16, 2, 4
    n_ls = level_resolutions(16, args.n_max, L)
    lat = build_lattice(n_ls)
    U, S = lat.num_nodes, lat.num_level_nodes
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    Ps = [2 ** 16, 2 ** 20, 2 ** 22] if args.quick else [2 ** 16, 2 ** 18, 2 ** 20, 2 ** 22, 2 ** 24]
    Ts = [2 ** 14, 2 ** 22] if args.quick else [2 ** 14, 2 ** 18, 2 ** 22]

    def timeit(fn, iters=20, warm=5):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.median(ts))

    g = torch.Generator(device="cuda").manual_seed(65535)
    for T in Ts:
        tables = [(torch.rand((T, F), device=dev, generator=g) * 2 - 1) * 1e-4 for _ in range(L)]
        tgrads = [torch.zeros((T, F), device=dev) for _ in range(L)]
        tab, gtab = _lib.make_tables(tables), _lib.make_tables(tgrads)
        utopi = torch.randint(0, T, (U, K), device=dev, generator=g, dtype=torch.int32)
        utopv = torch.rand((U, K), device=dev, generator=g).sort(dim=-1, descending=True).values
        nfeat = torch.empty((S, F), device=dev)
        dnf = torch.zeros((S, F), device=dev)
        dtv = torch.zeros((U, K), device=dev)
        node_fwd_b = S * (K * (8 + F * 4) + F * 4)
        node_bwd_b = S * (F * 4 + K * (8 + 2 * F * 4 + 4))
        t_node_f = timeit(lambda: _lib.call("gngf_node_features_fwd", lat, tab, T, F, K, 1, utopv.data_ptr(),
                                            utopi.data_ptr(), nfeat.data_ptr(), st))
        for P in Ps:
            x = torch.rand((P, 2), device=dev, generator=g)
            enc = torch.empty((P, L * F), device=dev)
            denc = torch.randn((P, L * F), device=dev, generator=g)
            pt_b = P * (8 + L * (4 * F * 4) + L * F * 4)
            res = {"P": P, "T": T, "L": L, "F": F, "K": K, "lattice_nodes": U, "level_nodes": S, "hbm_peak_gbs": hbm}
            t = timeit(lambda: _lib.call("gngf_encode_fwd", x.data_ptr(), P, lat, F, nfeat.data_ptr(), enc.data_ptr(),
                                         None, None, None, st))
            res["gngf_point_fwd"] = {"ms": t, "gbs": pt_b / t / 1e6, "frac": pt_b / t / 1e6 / hbm}
            res["gngf_node_fwd"] = {"ms": t_node_f, "gbs": node_fwd_b / t_node_f / 1e6}
            t = timeit(lambda: _lib.call("gngf_encode_bwd", x.data_ptr(), P, lat, F, denc.data_ptr(), dnf.data_ptr(), st))
            res["gngf_point_bwd"] = {"ms": t, "gbs": pt_b / t / 1e6, "frac": pt_b / t / 1e6 / hbm}
            t = timeit(lambda: _lib.call("gngf_node_features_bwd", lat, tab, gtab, T, F, K, 1, utopv.data_ptr(),
                                         utopi.data_ptr(), dnf.data_ptr(), dtv.data_ptr(), st))
            res["gngf_node_bwd"] = {"ms": t, "gbs": node_bwd_b / t / 1e6}
            t = timeit(lambda: _lib.call("gngf_encode_hash_fwd", x.data_ptr(), P, lat, tab, T, F, enc.data_ptr(), None,
                                         st))
            res["hash_fwd"] = {"ms": t, "gbs": pt_b / t / 1e6, "frac": pt_b / t / 1e6 / hbm}
            t = timeit(lambda: _lib.call("gngf_encode_hash_bwd", x.data_ptr(), P, lat, gtab, T, F, denc.data_ptr(), st))
            res["hash_bwd"] = {"ms": t, "gbs": pt_b / t / 1e6, "frac": pt_b / t / 1e6 / hbm}
            res["samples_per_s"] = {k: P / (res[k]["ms"] / 1e3) for k in ("gngf_point_fwd", "gngf_point_bwd", "hash_fwd",
                                                                         "hash_bwd")}
            print(json.dumps(res), flush=True)
            del x, enc, denc


if __name__ == "__main__":
    main()
