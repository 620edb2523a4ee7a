"""TEST INFRASTRUCTURE ONLY -- PyTorch-CPU restatement of the reference's hot path, used as the *timed* CPU
baseline (`bench.py --impl reference`, `cpu_baseline`).

The numpy oracle (gngf_oracle.py) is the parity checker; it is several times slower than the reference's own
CPU execution because numpy has no fused softmax / top-k / embedding kernels.  The reference itself is PyTorch
code that cannot travel to the GPU box, so this file restates its forward (models.py:394-484) and loss
(utils.py:91-174, functions.py:243-245) with the same ATen operations the reference issues (nn.functional.linear,
softmax, nan_to_num, topk, embedding lookups, autograd for the backward) and lets autograd differentiate it --
i.e. it costs what the reference costs on the same cores.  Pinned against the reference's golden vectors in
tests/test_oracle_golden.py.  Row-per-(point, level, corner) evaluation exactly as the reference: no lattice
de-duplication here.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def forward(params: dict, x: torch.Tensor, cfg: dict):
    """params: dict of lists of tensors (hpd_w, hpd_b, tables, mlp_w, mlp_b); returns rgb, probs, idx."""
    n_ls = torch.as_tensor(cfg["n_ls"], dtype=torch.int32).reshape(1, 1, -1, 1)
    cube = torch.tensor([[0, 1, 0, 1], [0, 0, 1, 1]], dtype=torch.int32).reshape(1, 2, 1, 4)
    with torch.no_grad():                                                        # models.py:486-502
        scaled = torch.mul(x.unsqueeze(-1).unsqueeze(-1), n_ls)
        grid = torch.add(torch.floor(scaled), cube)
    h = grid.permute(0, 2, 3, 1)                                                 # "p xy l v -> p l v xy"
    n = len(params["hpd_w"])
    for i, (w, b) in enumerate(zip(params["hpd_w"], params["hpd_b"])):           # models.py:105-106
        h = F.linear(h, w, b)
        h = torch.relu(h) if i < n - 1 else torch.softmax(h, dim=-1)
    probs = torch.nan_to_num(h)                                                  # models.py:111
    topv, topi = torch.topk(probs, k=cfg["topk_k"], dim=-1, largest=True, sorted=True)
    L = len(params["tables"])
    looked = torch.stack([F.embedding(topi[:, l], params["tables"][l]) for l in range(L)], dim=1)   # (P,L,4,K,F)
    feat = (looked * torch.softmax(topv, dim=-1).unsqueeze(-1)).sum(3)           # models.py:214-215
    a, d, s = grid[:, :, :, 0], grid[:, :, :, 3], scaled[:, :, :, 0]             # models.py:626-637
    coeffs = torch.stack([(d[:, 0] - s[:, 0]) * (d[:, 1] - s[:, 1]), (s[:, 0] - a[:, 0]) * (d[:, 1] - s[:, 1]),
                          (d[:, 0] - s[:, 0]) * (s[:, 1] - a[:, 1]), (s[:, 0] - a[:, 0]) * (s[:, 1] - a[:, 1])], dim=-1)
    enc = (feat * coeffs.unsqueeze(-1)).sum(2).reshape(x.shape[0], -1)           # (P, L*F) level-major
    hdn = enc
    m = len(params["mlp_w"])
    for i, (w, b) in enumerate(zip(params["mlp_w"], params["mlp_b"])):           # models.py:468-470
        hdn = F.linear(hdn, w, b)
        hdn = torch.relu(hdn) if i < m - 1 else torch.sigmoid(hdn)
    ret = topv if cfg.get("topk_only", False) else probs
    return hdn, ret, topi


def loss(rgb, target, probs, gamma, epsilon, l_mse, l_js_kl):
    """utils.py:91-174 + functions.py:243-245 (epoch 0: the collisions term is the scalar 1 per level)."""
    kl_fn = torch.nn.KLDivLoss(reduction="batchmean")
    N = probs.shape[-1]
    q = torch.ones(N) / float(N)
    levels = []
    for l in range(probs.shape[1]):
        p_out = probs[:, l, :].sum(0).sum(0) / (probs.shape[0] * probs.shape[2])
        kl = kl_fn(p_out.log(), q)
        m = (p_out + q) / 2
        js = (kl_fn(p_out.log(), m) + kl_fn(q.log(), m)) / 2
        levels.append(-(gamma + epsilon) * js + epsilon * kl)
    levels = torch.stack(levels)
    mse = F.mse_loss(rgb, target)
    return l_mse * mse + ((l_js_kl * levels) + 1).sum(0), mse, levels


def make_params(w: dict, seed: int):
    """Random-init parameters of the workload's architecture (nn.Linear-style bounds, tables U(-1e-4, 1e-4))."""
    g = torch.Generator().manual_seed(seed)

    def lin(i, o):
        b = 1 / np.sqrt(i)
        return ((torch.rand(o, i, generator=g) * 2 - 1) * b).requires_grad_(), \
               ((torch.rand(o, generator=g) * 2 - 1) * b).requires_grad_()

    widths = [2, *w["hpd"], w["T"]]
    mw = [w["L"] * w["F"], *w["mlp"], 3]
    hp = [lin(widths[i], widths[i + 1]) for i in range(len(widths) - 1)]
    ml = [lin(mw[i], mw[i + 1]) for i in range(len(mw) - 1)]
    tables = [((torch.rand(w["T"], w["F"], generator=g) * 2 - 1) * 1e-4).requires_grad_() for _ in range(w["L"])]
    return {"hpd_w": [a for a, _ in hp], "hpd_b": [b for _, b in hp], "tables": tables,
            "mlp_w": [a for a, _ in ml], "mlp_b": [b for _, b in ml]}


def make_step(w: dict, x: np.ndarray, y: np.ndarray, n_ls, seed: int = 65535):
    """Returns step() = forward + loss + backward + Adam (functions.py:96-127 groups) on CPU."""
    params = make_params(w, seed)
    cfg = {"n_ls": n_ls, "topk_k": w["K"], "topk_only": w["topk_only"]}
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    opt = torch.optim.Adam(
        [{"params": params["tables"], "lr": w["lr"]["encoding"], "weight_decay": w["wd"]["encoding"]},
         {"params": params["hpd_w"] + params["hpd_b"], "lr": w["lr"]["hpd"], "weight_decay": w["wd"]["hpd"]},
         {"params": params["mlp_w"] + params["mlp_b"], "lr": w["lr"]["mlp"], "weight_decay": w["wd"]["mlp"]}],
        betas=(0.9, 0.99), eps=1e-15)

    def step():
        opt.zero_grad()
        rgb, probs, _ = forward(params, xt, cfg)
        total, _, _ = loss(rgb, yt, probs, w["gamma"], w["epsilon"], w["l_mse"], w["l_js_kl"])
        total.backward()
        opt.step()
        return float(total)

    return step
