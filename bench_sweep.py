#!/usr/bin/env python
"""Encoding-only micro-bench sweep (BASELINE.json configs[4]): batch 2^16..2^24 points x table size 2^14..2^22,
L = 16 levels x F = 2 features, GNGF (HPD + top-k + gather / scatter) vs. the plain hash-mod baseline, forward and
backward separately.

    python bench_sweep.py [--quick] [--n-max 2048] > profiles/rNN_cfg5_sweep.jsonl

Per (P, T), with CUDA events around every C-ABI call (median over the iterations, an L2 flush before each):

  gngf  the drop-in module's own forward / backward (models.py:394-484 and its autograd) with the decoder, the loss and
        the int64 index output left out of the sums:
          hpd     the index SOURCE being compared with the hash: touched-node list, HPD layers, bf16 planes, streaming
                  softmax + top-k (forward); streaming backward + layer gradients (backward)
          encode  node pass (K = 4 table rows mixed per level node / table scatter-add + top-k adjoint) + point pass
                  (4 node-feature gathers per level / 4 vector reductions per level)
  hash  encode_hash_fwd (4 table gathers per level at (x ^ y*2654435761) mod T), encode_hash_bwd

The HPD evaluates the touched lattice nodes against all T slots: its cost is ~ nodes x T x 128 x 2 FLOP x 15 split
products (forward + backward); cells where that exceeds --hpd-cap FLOP are run WITHOUT the HPD arm and say so
("hpd": "capped") -- the corner where a step takes tens of seconds.

Peaks: encode kernels are put against the measured HBM copy peak (MEASURED_PEAKS.json) and, where the node-feature
array / tables are L2-resident, against the measured L2-resident copy bandwidth (measure_l2_copy below: the first line
this script prints).  Algorithmic bytes per point (hash-mod rows): fwd 8 + L*4*F*4 + L*F*4 = 648 B,
bwd 128 + 512 + 8 = 648 B; the GNGF point passes move the same bytes (node features instead of table rows) and the node
passes add S*(K*(8+F*4)+F*4) B fwd / S*(F*4+K*(8+2*F*4+4)) B bwd over the touched level nodes.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HPD_CALLS = ("gngf_lattice_mark_nodes", "gngf_compact_nodes", "gngf_hpd_first_layer_fwd_nodes",
             "gngf_hpd_first_layer_bwd_nodes", "gngf_linear_fwd", "gngf_linear_bwd", "gngf_split_bf16x3", "gngf_split_f16x2",
             "gngf_hpd_stream_fwd", "gngf_hpd_stream_fwd_refined", "gngf_hpd_stream_bwd", "gngf_hpd_stream_bwd_nodes",
             "gngf_scatter_node_rows", "gngf_softmax_topk_fwd", "gngf_hpd_dlogits", "gngf_tc_gemm_bf16x3",
             "gngf_hpd_small_fwd", "gngf_hpd_small_bwd", "gngf_hpd_small_fwd_enc", "gngf_hpd_small_bwd_enc")
ENCODE_CALLS = ("gngf_node_features_fwd", "gngf_node_features_bwd", "gngf_encode_fwd", "gngf_encode_bwd",
                "gngf_cell_to_node_counts")


def measure_l2_copy(torch, dev):
    """Best read+write GB/s of b.copy_(a) for buffers that stay L2-resident (the method MEASURED_PEAKS.json uses for
    HBM, at sizes below the 126 MB L2)."""
    best = {}
    for mb in (8, 16, 24, 32, 48):
        n = mb << 20
        a = torch.empty(n, dtype=torch.uint8, device=dev)
        b = torch.empty(n, dtype=torch.uint8, device=dev)
        for _ in range(5):
            b.copy_(a)
        ts = []
        for _ in range(30):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            b.copy_(a)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        best[mb] = 2 * n / (min(ts) * 1e-3) / 1e9
        del a, b
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--n-max", type=int, default=2048)
    ap.add_argument("--hpd-cap", type=float, default=6e14, help="executed FLOP (fwd + bwd) above which the HPD arm is skipped")
    ap.add_argument("--iters", type=int, default=7)
    args = ap.parse_args()
    import torch

    from collision_handling_in_instantngp_b200 import _lib, ops
    from collision_handling_in_instantngp_b200.models import GeneralNeuralGaugeFields

    dev = torch.device("cuda")
    L, F, K = 16, 2, 4
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = peaks.get("hbm_gbs", 6650.0)
    l2 = measure_l2_copy(torch, dev)
    l2_peak = max(l2.values())
    print(json.dumps({"l2_resident_copy_gbs": l2, "l2_peak_gbs": l2_peak, "hbm_peak_gbs": hbm,
                      "how": "torch b.copy_(a), read+write bytes, best of 30, buffers of 8..48 MB (L2 = 126 MB)"}), flush=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    Ps = [2 ** 16, 2 ** 20, 2 ** 22] if args.quick else [2 ** 16, 2 ** 18, 2 ** 20, 2 ** 22, 2 ** 24]
    Ts = [2 ** 14, 2 ** 22] if args.quick else [2 ** 14, 2 ** 16, 2 ** 18, 2 ** 20, 2 ** 22]

    class Prof:
        def __init__(self):
            self.ev = []

        def record(self, name, fn, a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*a)
            e1.record()
            self.ev.append((name, e0, e1))
            return rc

        def sums(self):
            torch.cuda.synchronize()
            out = {}
            for name, e0, e1 in self.ev:
                out[name] = out.get(name, 0.0) + e0.elapsed_time(e1)
            return out

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
        torch.manual_seed(65535)
        net = GeneralNeuralGaugeFields(input_dim=2, hash_table_size=T, num_levels=L, n_min=16, n_max=args.n_max,
                                       MLP_hidden_layers_widths=[64, 64], HPD_hidden_layers_widths=[32, 64, 128],
                                       HPD_out_features=T, feature_dim=F, topk_k=K, should_keep_topk_only=True)
        net.set_coord_bounds((0.0, 0.0), (1.0, 1.0))
        tables = net.encoding.tables()
        tgrads = [torch.zeros((T, F), device=dev) for _ in range(L)]
        tab, gtab = _lib.make_tables(tables), _lib.make_tables(tgrads)
        for P in Ps:
            x = torch.rand((P, 2), device=dev, generator=g)
            lat = net._lattice_for(x)
            U, S = lat.num_nodes, lat.num_level_nodes
            res = {"P": P, "T": T, "L": L, "F": F, "K": K, "n_max": args.n_max, "lattice_nodes": U, "level_nodes": S,
                   "hbm_peak_gbs": hbm, "l2_peak_gbs": l2_peak}
            pt_b = P * (8 + L * (4 * F * 4) + L * F * 4)
            # ---- hash-mod baseline (the index is computed in registers) ----
            enc = torch.empty((P, L * F), device=dev)
            denc = torch.randn((P, L * F), device=dev, generator=g)
            t = timeit(lambda: _lib.call("gngf_encode_hash_fwd", x.data_ptr(), P, lat, tab, T, F, enc.data_ptr(), None, st))
            res["hash_fwd"] = {"ms": t, "gbs": pt_b / t / 1e6, "frac_hbm": pt_b / t / 1e6 / hbm}
            t = timeit(lambda: _lib.call("gngf_encode_hash_bwd", x.data_ptr(), P, lat, gtab, T, F, denc.data_ptr(), st))
            res["hash_bwd"] = {"ms": t, "gbs": pt_b / t / 1e6, "frac_hbm": pt_b / t / 1e6 / hbm}
            del enc, denc
            # ---- GNGF: the module's forward / backward, per-call device times ----
            active = ops.active_nodes(x, lat).shape[0] if U >= ops.ACTIVE_NODES_MIN else U
            hpd_flop = 15 * 2.0 * active * T * 128
            res["hpd_rows"] = int(active)
            with_hpd = hpd_flop <= args.hpd_cap
            iters = args.iters if with_hpd else 0
            fwd, bwd = [], []
            if with_hpd:
                for it in range(iters + 2):
                    net.zero_grad(set_to_none=True)
                    flush.zero_()
                    pf = Prof()
                    _lib.PROFILER = pf
                    ops.CONCURRENT = False
                    rgb, probs, idx, _ = net(x, 1.0)
                    f = pf.sums()
                    pb = Prof()
                    _lib.PROFILER = pb
                    (rgb.sum() + probs.colsum.sum()).backward()
                    b = pb.sums()
                    _lib.PROFILER = None
                    ops.CONCURRENT = True
                    if it >= 2:
                        fwd.append(f)
                        bwd.append(b)
                    del rgb, probs, idx

                def med(rows, names):
                    return float(np.median([sum(r.get(n, 0.0) for n in names) for r in rows]))

                res["gngf_fwd"] = {"hpd_ms": med(fwd, HPD_CALLS), "encode_ms": med(fwd, ENCODE_CALLS),
                                   "point_ms": med(fwd, ("gngf_encode_fwd",)), "node_ms": med(fwd, ("gngf_node_features_fwd",))}
                res["gngf_bwd"] = {"hpd_ms": med(bwd, HPD_CALLS), "encode_ms": med(bwd, ENCODE_CALLS),
                                   "point_ms": med(bwd, ("gngf_encode_bwd",)), "node_ms": med(bwd, ("gngf_node_features_bwd",))}
                for k in ("gngf_fwd", "gngf_bwd"):
                    r = res[k]
                    r["ms"] = r["hpd_ms"] + r["encode_ms"]
                    r["point_gbs"] = pt_b / r["point_ms"] / 1e6
                    r["point_frac_hbm"] = r["point_gbs"] / hbm
                    r["point_frac_l2"] = r["point_gbs"] / l2_peak
                res["hpd_executed_tflops"] = {"fwd": 3 * 2.0 * active * T * 128 / (res["gngf_fwd"]["hpd_ms"] * 1e9),
                                              "bwd": 12 * 2.0 * active * T * 128 / (res["gngf_bwd"]["hpd_ms"] * 1e9)}
            else:
                # HPD arm capped: the gather / scatter passes alone, on random selections
                res["hpd"] = f"capped ({hpd_flop:.1e} executed FLOP per step > {args.hpd_cap:.0e})"
                utopi = torch.randint(0, T, (U, K), device=dev, generator=g, dtype=torch.int32)
                utopv = torch.rand((U, K), device=dev, generator=g).sort(dim=-1, descending=True).values
                nfeat = torch.empty((S, F), device=dev)
                dnf = torch.zeros((S, F), device=dev)
                dtv = torch.zeros((U, K), device=dev)
                enc = torch.empty((P, L * F), device=dev)
                denc = torch.randn((P, L * F), device=dev, generator=g)
                tn = timeit(lambda: _lib.call("gngf_node_features_fwd", lat, tab, T, F, K, 1, utopv.data_ptr(),
                                              utopi.data_ptr(), nfeat.data_ptr(), st), iters=7, warm=2)
                tp = timeit(lambda: _lib.call("gngf_encode_fwd", x.data_ptr(), P, lat, F, nfeat.data_ptr(), enc.data_ptr(),
                                              None, None, None, st), iters=7, warm=2)
                res["gngf_fwd"] = {"encode_ms": tn + tp, "point_ms": tp, "node_ms": tn, "point_gbs": pt_b / tp / 1e6,
                                   "point_frac_hbm": pt_b / tp / 1e6 / hbm, "point_frac_l2": pt_b / tp / 1e6 / l2_peak}
                tp = timeit(lambda: _lib.call("gngf_encode_bwd", x.data_ptr(), P, lat, F, denc.data_ptr(), dnf.data_ptr(),
                                              st), iters=7, warm=2)
                tn = timeit(lambda: _lib.call("gngf_node_features_bwd", lat, tab, gtab, T, F, K, 1, utopv.data_ptr(),
                                              utopi.data_ptr(), dnf.data_ptr(), dtv.data_ptr(), st), iters=7, warm=2)
                res["gngf_bwd"] = {"encode_ms": tn + tp, "point_ms": tp, "node_ms": tn, "point_gbs": pt_b / tp / 1e6,
                                   "point_frac_hbm": pt_b / tp / 1e6 / hbm, "point_frac_l2": pt_b / tp / 1e6 / l2_peak}
                del utopi, utopv, nfeat, dnf, dtv, enc, denc
            res["samples_per_s"] = {k: P / (res[k]["ms"] / 1e3) for k in ("gngf_fwd", "gngf_bwd", "hash_fwd", "hash_bwd")
                                    if "ms" in res[k]}
            print(json.dumps(res), flush=True)
            del x
            torch.cuda.empty_cache()
        del net, tables, tgrads
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
