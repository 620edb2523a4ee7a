"""TEST INFRASTRUCTURE ONLY -- CPU (numpy) restatement of the reference's GNGF training hot path.

This file is the parity oracle for the sm_100a CUDA path.  It is imported only by ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs, and only as the
checker (or as the timed CPU port) -- never by the product package, which fails loudly when its CUDA library
is missing.

Parity pinning: the reference has no tests or golden vectors of its own, so this oracle
is pinned against *outputs of the reference itself*: ``oracle/make_goldens.py`` imports the unmodified
reference from /root/reference (CPU, seed 65535) and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function below against those files.

Every function cites the reference lines it restates (paths relative to the reference root).  All
arithmetic is flat-index numpy in the dtype of the inputs (float32 = the reference's precision; float64 is
used by the tests to budget rounding error).

Layout conventions (identical to the reference):
    x        (P, 2)            input coordinates, column 0 = "x", column 1 = "y"
    n_ls     (L,) int32        per-level resolutions
    corner v in {0,1,2,3} <-> (dx, dy) = (0,0), (1,0), (0,1), (1,1)          (models.py:322-331)
    rows     (P, L, 4)         one HPD row per (point, level, corner)
    enc      (P, L*F)          level-major, feature inner                     (models.py:651)
"""
from __future__ import annotations

import numpy as np

CORNER_DX = np.array([0, 1, 0, 1])
CORNER_DY = np.array([0, 0, 1, 1])
PRIMES = (1, 2654435761, 805459861)  # models.py:346


# --------------------------------------------------------------------------------------------------------
# a-1  level table                                                                     models.py:305-317
# --------------------------------------------------------------------------------------------------------
def level_resolutions(n_min: int, n_max: int, num_levels: int) -> np.ndarray:
    """n_l = floor(n_min * b**l), b = exp((ln n_max - ln n_min)/(L-1)), all in numpy float64."""
    b = np.exp((np.log(n_max) - np.log(n_min)) / (num_levels - 1))
    return np.array([np.floor(n_min * b ** l) for l in range(num_levels)]).astype(np.int32)


# --------------------------------------------------------------------------------------------------------
# a-2  _scale_to_grid                                                                  models.py:486-502
# --------------------------------------------------------------------------------------------------------
def scale_to_grid(x: np.ndarray, n_ls: np.ndarray):
    """scaled (P,2,L,1) = x * n_l (one rounding, in x.dtype); grid (P,2,L,4) = floor(scaled) + hypercube."""
    dt = x.dtype
    scaled = x[:, :, None, None] * n_ls.astype(dt)[None, None, :, None]
    cube = np.stack([CORNER_DX, CORNER_DY]).astype(dt)[None, :, None, :]          # (1,2,1,4)
    grid = np.floor(scaled) + cube
    return scaled, grid


# --------------------------------------------------------------------------------------------------------
# a-7  _fast_hash (hash-function mode)                                                 models.py:504-528
# --------------------------------------------------------------------------------------------------------
def fast_hash(grid_int: np.ndarray, table_size: int) -> np.ndarray:
    """grid_int (P,2,L,4) int32 -> (P,L,4) int64.

    `grid[:, i] * prime_i` multiplies an int32 tensor by a 0-dim int64 tensor: the result stays int32 and
    wraps; the xor with the int64 accumulator sign-extends; torch.remainder is the non-negative modulo.
    """
    acc = np.zeros(grid_int[:, 0].shape, dtype=np.int64)
    for i in range(grid_int.shape[1]):
        prod = (grid_int[:, i].astype(np.int64) * np.int64(PRIMES[i])).astype(np.int32)   # wrap to int32
        acc = np.bitwise_xor(prod.astype(np.int64), acc)
    return np.mod(acc, np.int64(table_size))


# --------------------------------------------------------------------------------------------------------
# a-4/a-5  HashProbDistribution MLP + softmax + nan_to_num                    models.py:80-88, 105-111
# --------------------------------------------------------------------------------------------------------
def softmax_lastdim(z: np.ndarray) -> np.ndarray:
    m = z.max(axis=-1, keepdims=True)
    e = np.exp(z - m)
    return e / e.sum(axis=-1, keepdims=True)


def hpd_mlp(inp: np.ndarray, weights, biases):
    """inp (...,in) -> list of activations [inp, h1, ..., h_last] and logits.  ReLU between layers."""
    acts = [inp]
    h = inp
    for i, (w, b) in enumerate(zip(weights, biases)):
        z = h @ w.T + b
        if i < len(weights) - 1:
            h = np.maximum(z, 0)
            acts.append(h)
    return acts, z


# --------------------------------------------------------------------------------------------------------
# a-6  DifferentiableTopk.forward                                                        models.py:7-19
# --------------------------------------------------------------------------------------------------------
def topk_sorted(p: np.ndarray, k: int):
    """largest-k along the last dim, sorted descending; ties -> lowest index first (torch leaves the tie
    order unspecified; the CUDA kernel and this oracle both define it this way)."""
    order = np.argsort(-p, axis=-1, kind="stable")[..., :k]
    return np.take_along_axis(p, order, axis=-1), order.astype(np.int64)


def hpd_forward(inp: np.ndarray, weights, biases, k: int, force_idx=None):
    """HashProbDistribution.forward (models.py:90-123): returns probs, topk_probs, topk_idx, acts, logits.
    force_idx: use this selection instead of running top-k (error budgeting in float64 must differentiate the
    same discrete selection as the float32 run, near-ties included)."""
    acts, logits = hpd_mlp(inp, weights, biases)
    probs = softmax_lastdim(logits)
    probs = np.nan_to_num(probs)                                               # models.py:111
    if force_idx is None:
        topv, topi = topk_sorted(probs, k)
    else:
        topi = force_idx
        topv = np.take_along_axis(probs, topi, axis=-1)
    return probs, topv, topi, acts, logits


# --------------------------------------------------------------------------------------------------------
# a-8/a-9  MultiResHashEncoding.forward                                              models.py:173-229
# --------------------------------------------------------------------------------------------------------
def mix_weights(topv: np.ndarray, mode):
    """mode True: softmax over K of the top-k *probabilities* (models.py:214-215);
    mode False: topv / sum(topv) (models.py:216-217); mode None: raw topv (models.py:212-213)."""
    if mode is None:
        return topv
    if mode:
        return softmax_lastdim(topv)
    return topv / topv.sum(axis=-1, keepdims=True)


def encoding_forward(tables, idx: np.ndarray, topv: np.ndarray, mode=True):
    """tables: list of L arrays (T,F); idx (P,L,4,K) int; topv (P,L,4,K) -> feat (P,F,L,4), w (P,L,4,K),
    gathered g (P,L,4,K,F)."""
    L = len(tables)
    g = np.stack([tables[l][idx[:, l]] for l in range(L)], axis=1)               # (P,L,4,K,F)
    w = mix_weights(topv, mode)
    if mode is False:
        feat = (g * topv[..., None]).sum(axis=3) / topv.sum(axis=-1)[..., None]  # reference's op order
    else:
        feat = (g * w[..., None]).sum(axis=3)                                    # (P,L,4,F)
    return np.transpose(feat, (0, 3, 1, 2)), w, g                                # "p l f v -> p f l v"


def encoding_forward_hash(tables, idx: np.ndarray):
    """hash-function mode (models.py:181-190): idx (P,L,4) -> feat (P,F,L,4)."""
    L = len(tables)
    g = np.stack([tables[l][idx[:, l]] for l in range(L)], axis=1)               # (P,L,4,F)
    return np.transpose(g, (0, 3, 1, 2))


# --------------------------------------------------------------------------------------------------------
# a-10  _bilinear_interpolate                                                        models.py:621-655
# --------------------------------------------------------------------------------------------------------
def bilinear_coeffs(scaled: np.ndarray, grid: np.ndarray) -> np.ndarray:
    """(P,L,4) weights in corner order [(xd-x)(yd-y), (x-xa)(yd-y), (xd-x)(y-ya), (x-xa)(y-ya)]."""
    a = grid[:, :, :, 0]          # (P,2,L)  floor corner
    d = grid[:, :, :, 3]          # (P,2,L)  floor+1 corner
    s = scaled[:, :, :, 0]
    return np.stack([
        (d[:, 0] - s[:, 0]) * (d[:, 1] - s[:, 1]),
        (s[:, 0] - a[:, 0]) * (d[:, 1] - s[:, 1]),
        (d[:, 0] - s[:, 0]) * (s[:, 1] - a[:, 1]),
        (s[:, 0] - a[:, 0]) * (s[:, 1] - a[:, 1]),
    ], axis=-1)


def bilinear_interpolate(scaled, grid, feat):
    """feat (P,F,L,4) -> enc (P, L*F)  ("p f l -> p (l f)": level-major, feature inner)."""
    wb = bilinear_coeffs(scaled, grid)                                           # (P,L,4)
    summed = (feat * wb[:, None]).sum(axis=-1)                                   # (P,F,L)
    P, F, L = summed.shape
    return np.transpose(summed, (0, 2, 1)).reshape(P, L * F), wb


# --------------------------------------------------------------------------------------------------------
# a-11  decoder MLP                                                         models.py:382-392, 468-470
# --------------------------------------------------------------------------------------------------------
def decoder_forward(enc, weights, biases, leaky=False):
    acts = [enc]
    h = enc
    n = len(weights)
    for i, (w, b) in enumerate(zip(weights, biases)):
        z = h @ w.T + b
        if i < n - 1:
            h = np.where(z > 0, z, z * z.dtype.type(0.01)) if leaky else np.maximum(z, 0)
        else:
            h = 1 / (1 + np.exp(-z))
        acts.append(h)
    return acts


# --------------------------------------------------------------------------------------------------------
# Loss                                                        utils.py:91-174, functions.py:243-245
# --------------------------------------------------------------------------------------------------------
def level_divergences(pbar: np.ndarray, gamma: float, epsilon: float):
    """pbar (L,N) mean slot distribution.  Returns (level_loss (L,), dloss/dpbar (L,N)).

    kl = sum q (ln q - ln pbar) / N                       (KLDivLoss 'batchmean' on a 1-D vector, utils.py:144)
    js = [sum m (ln m - ln pbar) + sum m (ln m - ln q)] / (2N),  m = (pbar+q)/2   (utils.py:167-168;
         the gradient also flows through the *target* m)
    level = -(gamma+epsilon) js + epsilon kl                                       (utils.py:127)
    """
    dt = pbar.dtype
    N = pbar.shape[-1]
    q = dt.type(1.0) / dt.type(N)
    lp = np.log(pbar)
    lq = np.log(q)
    kl = (q * (lq - lp)).sum(-1) / N
    m = (pbar + q) / 2
    lm = np.log(m)
    js = ((m * (lm - lp)).sum(-1) / N + (m * (lm - lq)).sum(-1) / N) / 2
    level = -(gamma + epsilon) * js + epsilon * kl
    dkl = -q / pbar / N
    djs = (0.5 * (lm - lp) + 0.5 - m / pbar + 0.5 * (lm - lq) + 0.5) / (2 * N)
    dlevel = -(gamma + epsilon) * djs + epsilon * dkl
    return level.astype(dt), dlevel.astype(dt), kl, js


def total_loss(out, target, pbar, gamma, epsilon, l_mse, l_js_kl, l_collisions, collisions_losses=None):
    """functions.py:243-245.  collisions_losses None == epoch 0 (empty tensors -> the scalar 1 per level)."""
    mse = np.mean((out - target) ** 2, dtype=out.dtype)
    level, dlevel, _, _ = level_divergences(pbar, gamma, epsilon)
    coll = np.ones_like(level) if collisions_losses is None else l_collisions * collisions_losses
    total = l_mse * mse + (l_js_kl * level + coll).sum()
    return total, mse, level, dlevel


# --------------------------------------------------------------------------------------------------------
# f-1  calc_hash_collisions                                                           models.py:568-619
# --------------------------------------------------------------------------------------------------------
def calc_hash_collisions(indices: np.ndarray, n_ls: np.ndarray, table_size: int):
    """indices (P,L,4,Kc) any dtype (train_step passes float32).  Returns (collisions (L,), min_possible (L,))."""
    L = len(n_ls)
    nodes = np.array([(int(n) + 1) ** 2 for n in n_ls], dtype=np.float64)
    per_col = np.empty((indices.shape[-1], L), dtype=np.float32)
    for k in range(indices.shape[-1]):
        for l in range(L):
            per_col[k, l] = nodes[l] - np.unique(indices[:, l, :, k].reshape(-1)).shape[0]
    coll = per_col.mean(axis=0)
    coll[coll < 0] = 0
    minp = nodes - table_size
    minp[minp < 0] = 0
    return coll, minp.astype(np.int64)


def calc_hash_collisions_hash_mode(indices: np.ndarray, n_ls: np.ndarray, table_size: int):
    """should_use_hash_function branch (models.py:574-585): indices (P,L,4); no clamp on `collisions`."""
    L = len(n_ls)
    nodes = np.array([(int(n) + 1) ** 2 for n in n_ls], dtype=np.int64)
    coll = np.array([nodes[l] - np.unique(indices[:, l].reshape(-1)).shape[0] for l in range(L)])
    minp = nodes - table_size
    minp[minp < 0] = 0
    return coll, minp


# --------------------------------------------------------------------------------------------------------
# f-4  _calc_counts_per_level                                                          models.py:530-566
# --------------------------------------------------------------------------------------------------------
def calc_counts_per_level(hashed: np.ndarray, grid: np.ndarray):
    """hashed (P,L,4) slot per (point, level, corner) [the best top-k column, or the hash]; grid (P,2,L,4) corners.
    Per level the rows "p (v xy)" (the 8 corner coordinates of a point's cell) are de-duplicated with
    np.unique(axis=0, return_index=True): `first[j]` is the index of the first POINT (batch order) in distinct cell j.
    The reference then uses these point indices to index the FLATTENED "(p v)" slot vector (models.py:556-558), so the
    slot it counts for cell j is that of corner (first[j] % 4) of point (first[j] // 4) -- restated as is.
    Returns a list (per level) of {slot: number of distinct cells counted for it}."""
    from collections import Counter
    P, _, L, _ = grid.shape
    rows = np.transpose(grid, (2, 0, 3, 1)).reshape(L, P, -1)          # "p xy l v -> l p (v xy)"
    flat = np.transpose(hashed, (1, 0, 2)).reshape(L, -1)             # "p l v -> l (p v)"
    out = []
    for l in range(L):
        _, first = np.unique(rows[l], axis=0, return_index=True)
        out.append(dict(Counter(flat[l][first].tolist())))
    return out


# --------------------------------------------------------------------------------------------------------
# whole path: forward                                                                models.py:394-484
# --------------------------------------------------------------------------------------------------------
def gngf_forward(params: dict, x: np.ndarray, cfg: dict) -> dict:
    """params: {'hpd_w': [...], 'hpd_b': [...], 'tables': [...], 'mlp_w': [...], 'mlp_b': [...]}
    cfg: {'n_ls', 'table_size', 'topk_k', 'mix_mode' (True/False/None), 'use_hash', 'leaky', 'topk_only'}."""
    n_ls = cfg["n_ls"]
    K = cfg.get("topk_k", 4)
    mode = cfg.get("mix_mode", True)
    out = {}
    scaled, grid = scale_to_grid(x, n_ls)
    out["scaled"], out["grid"] = scaled, grid
    if cfg.get("use_hash", False):
        idx = fast_hash(grid.astype(np.int32), cfg["table_size"])                # (P,L,4)
        out["idx"] = idx
        feat = encoding_forward_hash(params["tables"], idx)
    else:
        inp = np.transpose(grid, (0, 2, 3, 1))                                   # "p xy l v -> p l v xy"
        probs, topv, topi, acts, logits = hpd_forward(inp, params["hpd_w"], params["hpd_b"], K, cfg.get("force_idx"))
        out.update(hpd_in=inp, probs=probs, topv=topv, idx=topi, hpd_acts=acts, logits=logits)
        feat, w, g = encoding_forward(params["tables"], topi, topv, mode)
        out.update(mix_w=w, gathered=g)
    out["feat"] = feat
    enc, wb = bilinear_interpolate(scaled, grid, feat)
    out["enc"], out["wb"] = enc, wb
    acts = decoder_forward(enc, params["mlp_w"], params["mlp_b"], cfg.get("leaky", False))
    out["mlp_acts"] = acts
    out["rgb"] = acts[-1]
    if not cfg.get("use_hash", False):
        ret = out["topv"] if cfg.get("topk_only", False) else out["probs"]
        out["ret_probs"] = ret
        # utils.py:138 (p.sum(0).sum(0) / div); accumulated in float64 so that the oracle's own rounding
        # (numpy does not sum pairwise across non-contiguous axes) stays below the 1e-5 parity bar at P ~ 6e4
        out["pbar"] = (ret.sum(axis=(0, 2), dtype=np.float64) / (ret.shape[0] * ret.shape[2])).astype(ret.dtype)
    return out


# --------------------------------------------------------------------------------------------------------
# whole path: backward (closed form of what autograd does for a-4 ... a-11 + Loss)   SURVEY.md 8a-14
# --------------------------------------------------------------------------------------------------------
def _mlp_backward(acts, weights, dz_last, relu_like=True, leaky=False, masks=None):
    """acts[i] is the input of layer i; dz_last is the grad wrt the last layer's pre-activation.
    Returns (dW list, db list, dX of layer 0).  masks (optional): masks[i] (rows, width_i) booleans replacing
    `acts[i] > 0` for i >= 1 -- differentiates a GIVEN ReLU pattern (a forward whose pre-activations differ from this
    one's at rounding level sits on the other side of a kink for a few units; the derivative is discontinuous there)."""
    n = len(weights)
    dws, dbs = [None] * n, [None] * n
    dz = dz_last
    for i in range(n - 1, -1, -1):
        a = acts[i].reshape(-1, acts[i].shape[-1])
        dz2 = dz.reshape(-1, dz.shape[-1])
        dws[i] = dz2.T @ a
        dbs[i] = dz2.sum(axis=0)
        dx = dz2 @ weights[i]
        if i > 0:
            pos = (a > 0) if masks is None or masks[i] is None else masks[i].reshape(a.shape)
            if leaky:
                dz = np.where(pos, dx, dx * dx.dtype.type(0.01))
            else:
                dz = dx * pos
        else:
            dz = dx
    return dws, dbs, dz


def gngf_backward(params: dict, x: np.ndarray, target: np.ndarray, cfg: dict, fwd: dict, loss_cfg: dict) -> dict:
    """Gradients of total_loss wrt every parameter.  loss_cfg: gamma, epsilon, l_mse, l_js_kl."""
    dt = x.dtype
    P = x.shape[0]
    L = len(cfg["n_ls"])
    mode = cfg.get("mix_mode", True)
    use_hash = cfg.get("use_hash", False)
    rgb = fwd["rgb"]
    grads = {}

    # MSE -> sigmoid
    dout = loss_cfg["l_mse"] * 2 * (rgb - target) / dt.type(rgb.size)
    dz = dout * rgb * (1 - rgb)
    # cfg["force_mlp_masks"]: [None, (P,w1) bool, (P,w2) bool] -- differentiate a given ReLU pattern (see _mlp_backward)
    dws, dbs, denc = _mlp_backward(fwd["mlp_acts"][:-1], params["mlp_w"], dz, leaky=cfg.get("leaky", False),
                                   masks=cfg.get("force_mlp_masks"))
    grads["mlp_w"], grads["mlp_b"] = dws, dbs

    F = params["tables"][0].shape[1]
    denc = denc.reshape(P, L, F)
    dfeat = denc[:, :, None, :] * fwd["wb"][..., None]                           # (P,L,4,F)
    tgrads = [np.zeros_like(t) for t in params["tables"]]
    if use_hash:
        for l in range(L):
            np.add.at(tgrads[l], fwd["idx"][:, l].reshape(-1), dfeat[:, l].reshape(-1, F))
        grads["tables"] = tgrads
        return grads

    w, g, topv, idx, probs = fwd["mix_w"], fwd["gathered"], fwd["topv"], fwd["idx"], fwd["probs"]
    dg = dfeat[:, :, :, None, :] * w[..., None]                                   # (P,L,4,K,F)
    for l in range(L):
        np.add.at(tgrads[l], idx[:, l].reshape(-1), dg[:, l].reshape(-1, F))
    grads["tables"] = tgrads
    dw = (dfeat[:, :, :, None, :] * g).sum(-1)                                    # (P,L,4,K)
    if mode is None:
        dtv = dw
    elif mode:
        dtv = w * (dw - (dw * w).sum(-1, keepdims=True))
    else:
        s = topv.sum(-1, keepdims=True)
        dtv = (dw - (dw * w).sum(-1, keepdims=True)) / s

    _, dlevel, _, _ = level_divergences(fwd["pbar"], loss_cfg["gamma"], loss_cfg["epsilon"])
    dpbar = loss_cfg["l_js_kl"] * dlevel / dt.type(4 * P)                        # (L,N) per-row share
    if cfg.get("topk_only", False):
        dtv = dtv + dpbar[None, :, None, :]
        G = np.zeros_like(probs)
    else:
        G = np.broadcast_to(dpbar[None, :, None, :], probs.shape).copy()
    if cfg.get("drop_topk_adjoint", False):
        # params.should_inplace_scatter is None (models.py:30-31): the scatter result is discarded, so the adjoint of
        # the selected values never reaches `probs`
        dtv = np.zeros_like(dtv)
    np.put_along_axis(G, idx, np.take_along_axis(G, idx, -1) + dtv, axis=-1)     # DifferentiableTopk.backward
    dlogit = probs * (G - (G * probs).sum(-1, keepdims=True))
    grads["dlogit"] = dlogit
    dws, dbs, _ = _mlp_backward(fwd["hpd_acts"], params["hpd_w"], dlogit)
    grads["hpd_w"], grads["hpd_b"] = dws, dbs
    return grads
