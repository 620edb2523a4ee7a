"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the build container (needs /root/reference; CPU only):

    python oracle/make_goldens.py            # all cases
    python oracle/make_goldens.py cfg2_small # one case

Each case seeds torch with 65535 (functions.py:43-47), constructs the reference's
``GeneralNeuralGaugeFields`` + ``Loss``, runs forward, the loss assembly of functions.py:243-245 and
``loss.backward()``, and stores inputs, parameters, selected intermediates, outputs and parameter
gradients.  The committed .npz files are what pins ``oracle/gngf_oracle.py`` (tests/test_oracle_golden.py)
and, through it, the CUDA path.  The reference tree is never copied; only its numeric outputs are stored.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")
SEED = 2 ** 16 - 1

BASE = dict(T=256, L=4, n_min=8, n_max=32, F=2, K=4, hpd=[32, 64, 128], mlp=[64, 64], P=333,
            topk_only=False, mix_mode=True, use_hash=False, leaky=False, bw=False,
            gamma=-2.0, epsilon=1.0, l_mse=1.0, l_js_kl=1.0, l_collisions=1e-3, with_collisions=False,
            coords="image", inplace=True, counts=False)

CASES = {
    # grid-search ID 4061 (README.md:15-18): the published best parameters
    "cfg2_small": dict(),
    "cfg2_topk_only": dict(topk_only=True),
    "cfg2_epoch1": dict(with_collisions=True),
    "mix_weighted_avg": dict(mix_mode=False, P=128),
    "mix_raw": dict(mix_mode=None, P=128),
    "hash_mode": dict(use_hash=True, P=257),
    "k1": dict(K=1, P=128),
    "k20": dict(K=20, P=96),
    "bw_leaky": dict(bw=True, leaky=True, P=128),
    "l16_t1024": dict(L=16, n_min=16, n_max=508, T=1024, P=64, coords="uniform"),
    "l8_t4096_topk_only": dict(L=8, n_min=16, n_max=339, T=4096, P=48, topk_only=True, coords="uniform"),
    "js_only": dict(gamma=-1.0, epsilon=0.0, P=64),     # should_sum_js_kl_div False, should_js_div True
    "kl_only": dict(gamma=-1.0, epsilon=1.0, P=64),     # should_sum_js_kl_div False, should_js_div False
    # params.should_inplace_scatter = None: DifferentiableTopk.backward discards its scatter (models.py:30-31)
    "scatter_none": dict(inplace=None, P=128),
    "scatter_none_topk_only": dict(inplace=None, topk_only=True, P=128),
    # should_calc_counts=True: the per-level histograms of _calc_counts_per_level (models.py:530-566)
    "counts": dict(counts=True, P=333),
    "counts_hash": dict(counts=True, use_hash=True, P=257),
    "counts_l16": dict(counts=True, L=16, n_min=16, n_max=508, T=1024, P=200, coords="uniform"),
}


def _image_coords(ref):
    """main.py:42-51: (row, col) / (max(w,h)-1), plus the rgb targets."""
    import cv2
    img = cv2.cvtColor(cv2.imread(os.path.join(ref_shim.REFERENCE_DIR, "images", "strawberry.jpeg"))[:, :, :3],
                       cv2.COLOR_BGR2RGB)
    h, w = img.shape[:2]
    X = np.stack(np.meshgrid(range(h), range(w), indexing="ij"), axis=-1).reshape(-1, 2)
    x = torch.tensor(X).float() / (max(w, h) - 1)
    y = torch.tensor(img.reshape(-1, 3) / 255).float()
    return x, y, h, w, img


def run_case(ref, name, over):
    c = dict(BASE)
    c.update(over)
    ref_shim.set_flag(ref, "should_use_hash_function", c["use_hash"])
    ref_shim.set_flag(ref, "should_softmax_topk_features", c["mix_mode"])
    ref_shim.set_flag(ref, "should_leaky_relu", c["leaky"])
    ref_shim.set_flag(ref, "should_inplace_scatter", c["inplace"])
    torch.manual_seed(SEED)
    net = ref.models.GeneralNeuralGaugeFields(
        input_dim=2, hash_table_size=c["T"], num_levels=c["L"], n_min=c["n_min"], n_max=c["n_max"],
        MLP_hidden_layers_widths=c["mlp"], HPD_hidden_layers_widths=c["hpd"], HPD_out_features=c["T"],
        feature_dim=c["F"], topk_k=c["K"], should_keep_topk_only=c["topk_only"], should_bw=c["bw"])
    # the reference initialises tables with U(-1e-4, 1e-4); after a few optimizer steps they are O(1e-2).
    # Scale them up so that gradients through the tables are not lost in fp32 noise in the comparison.
    with torch.no_grad():
        for t in net.encoding._hash_tables:
            t.weight.mul_(300.0)
    C = 1 if c["bw"] else 3
    g = torch.Generator().manual_seed(SEED)
    if c["coords"] == "image":
        x_all, y_all, h, w, _ = _image_coords(ref)
        perm = torch.randperm(x_all.shape[0], generator=g)[: c["P"] - 3]
        # always include the last image row (x == 1.0 exactly -> corner n_l + 1 with weight 0) and the origin
        extra = torch.tensor([0, x_all.shape[0] - 1, x_all.shape[0] - w])
        sel = torch.cat([perm, extra])
        x, y = x_all[sel], y_all[sel][:, :C]
    else:
        x = torch.rand(c["P"], 2, generator=g)
        y = torch.rand(c["P"], C, generator=g)
    loss_fn = ref.utils.Loss(delta=1, gamma=c["gamma"], epsilon=c["epsilon"])
    if c["with_collisions"]:
        coll = torch.tensor([0.0, 20.0, 200.0, 850.0])[: c["L"]]
        minp = torch.tensor([0.0, 0.0, 185.0, 833.0])[: c["L"]]
    else:
        coll, minp = torch.tensor([]), torch.tensor([])

    rgb, probs, idx, counts = net(x, 1.0, should_calc_counts=c["counts"])
    if c["use_hash"]:
        mse, kl, coll_l = loss_fn(rgb, y, None, None, None, None)
        loss = c["l_mse"] * mse
    else:
        mse, kl, coll_l = loss_fn(rgb, y, probs.shape[-1], probs, coll, minp)
        # functions.py:243-245
        loss = c["l_mse"] * mse
        loss = loss + ((c["l_js_kl"] * kl) + (c["l_collisions"] * coll_l if coll_l.nelement() != 0 else 1)).sum(0)
    loss.backward()

    with torch.no_grad():
        scaled, grid = net._scale_to_grid(x)
    out = {"x": x, "y": y, "n_ls": net._n_ls.flatten(), "scaled": scaled, "grid": grid,
           "rgb": rgb, "idx": idx, "mse": mse, "loss": loss}
    if not c["use_hash"]:
        out.update(kl_levels=kl, ret_probs_head=probs[:4],
                   pbar=probs.sum(0).sum(1) / (probs.shape[0] * probs.shape[2]))
        if coll_l.nelement() != 0:
            out.update(collisions=coll, min_possible=minp, coll_losses=coll_l)
        # intermediates recomputed through the reference's own sub-modules
        with torch.no_grad():
            inp = grid.permute(0, 2, 3, 1)
            full, topv, topi = net.HPD(inp)
            feat = net.encoding(topi, topv)
            enc = net._bilinear_interpolate(scaled, grid, feat)
        assert torch.equal(topi, idx)
        out.update(topv=topv, feat=feat, enc=enc, probs_head=full[:4])
        coll_k, minp_k = net.calc_hash_collisions(idx.float())
        out.update(chc_collisions=coll_k, chc_min_possible=minp_k)
    else:
        with torch.no_grad():
            feat = net.encoding(idx, None)
            enc = net._bilinear_interpolate(scaled, grid, feat)
        out.update(feat=feat, enc=enc)
        coll_k, minp_k = net.calc_hash_collisions(idx)
        out.update(chc_collisions=coll_k, chc_min_possible=minp_k)
    if c["counts"]:                    # list (per level) of {slot: number of distinct grid corners hashed to it}
        assert len(counts) == c["L"]
        for l, d in enumerate(counts):
            keys = np.array(sorted(d), dtype=np.int64)
            out[f"counts_keys_{l}"] = keys
            out[f"counts_vals_{l}"] = np.array([d[k] for k in keys], dtype=np.int64)
    for k, v in net.state_dict().items():
        if k.startswith("_batch_norm"):
            continue
        out["param." + k] = v
    for k, v in net.named_parameters():
        if k.startswith("_batch_norm") or v.grad is None:
            continue
        out["grad." + k] = v.grad
    arrays = {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in out.items()}
    arrays["cfg"] = np.array(repr(c))
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **arrays)
    print(f"{name}: loss={float(loss):.8f} mse={float(mse):.8f} P={x.shape[0]}")


def level_tables(ref):
    """a-1: n_l for a spread of (n_min, n_max, L), through the reference constructor."""
    rows = []
    for n_min, n_max, L in [(8, 32, 4), (16, 508, 16), (16, 8192, 16), (16, 339, 8), (2, 1024, 10), (16, 2048, 16),
                            (4, 64, 5), (16, 512, 16), (8, 4096, 12)]:
        torch.manual_seed(SEED)
        net = ref.models.GeneralNeuralGaugeFields(2, 16, L, n_min, n_max, [8], [8], HPD_out_features=16, topk_k=1)
        rows.append((n_min, n_max, L, net._n_ls.flatten().numpy()))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "level_tables.npz"),
                        keys=np.array([r[:3] for r in rows]), **{f"n_ls_{i}": r[3] for i, r in enumerate(rows)})
    print("level_tables:", [list(r[3][-2:]) for r in rows])


def strawberry_pixels(ref):
    """The decoded strawberry image (uint8) -- the PSNR parity target; the GPU box has no reference tree."""
    _, _, h, w, img = _image_coords(ref)
    np.savez_compressed(os.path.join(GOLDEN_DIR, "strawberry_u8.npz"), img=img)
    print("strawberry:", img.shape)


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    ref = ref_shim.load_reference("cpu")
    want = sys.argv[1:] or list(CASES) + ["level_tables", "strawberry"]
    for name in want:
        if name == "level_tables":
            level_tables(ref)
        elif name == "strawberry":
            strawberry_pixels(ref)
        else:
            run_case(ref, name, CASES[name])


if __name__ == "__main__":
    main()
