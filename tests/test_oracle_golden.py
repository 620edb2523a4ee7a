"""CPU: pins oracle/gngf_oracle.py against golden vectors produced by the unmodified reference
(oracle/make_goldens.py).  Tolerances: integer/index outputs exact; fp32 forward 1e-5 relative; parameter
gradients 1e-4 relative (BASELINE.json north_star)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import gngf_oracle as O  # noqa: E402
from golden_util import (ALL_CASES, COUNTS_CASES, GNGF_CASES, GOLDEN_DIR, golden_counts, load, loss_cfg, oracle_cfg,  # noqa: E402
                         params_of, rel_err)

FWD_TOL = 1e-5
GRAD_TOL = 1e-4


def test_level_tables_match_reference():
    z = np.load(os.path.join(GOLDEN_DIR, "level_tables.npz"))
    for i, (n_min, n_max, L) in enumerate(z["keys"]):
        got = O.level_resolutions(int(n_min), int(n_max), int(L))
        assert got.dtype == np.int32
        np.testing.assert_array_equal(got, z[f"n_ls_{i}"])
    # the two documented surprises
    assert O.level_resolutions(16, 8192, 16)[-1] == 8191
    assert O.level_resolutions(16, 339, 8)[-1] == 338


@pytest.mark.parametrize("name", ALL_CASES)
def test_scale_to_grid_bit_exact(name):
    g = load(name)
    scaled, grid = O.scale_to_grid(g["x"], g["n_ls"])
    assert scaled.dtype == np.float32
    np.testing.assert_array_equal(scaled, g["scaled"])
    np.testing.assert_array_equal(grid, g["grid"])
    # integer corners never exceed n_l + 1 and x == 1.0 hits exactly n_l
    assert (grid >= 0).all() and (grid[:, :, :, 3] <= g["n_ls"][None, None, :] + 1).all()


def test_fast_hash_bit_exact():
    g = load("hash_mode")
    _, grid = O.scale_to_grid(g["x"], g["n_ls"])
    idx = O.fast_hash(grid.astype(np.int32), g["cfg"]["T"])
    assert idx.dtype == np.int64
    np.testing.assert_array_equal(idx, g["idx"])
    # larger coordinates exercise the int32 wrap of y * 2654435761
    big = np.stack([np.arange(0, 9000, 7), np.arange(9000, 0, -7)], 1)[:, :, None, None].astype(np.int32)
    big = np.broadcast_to(big, (big.shape[0], 2, 1, 4))
    h = O.fast_hash(big, 2 ** 19)
    ux = big[:, 0].astype(np.uint32)
    uy = (big[:, 1].astype(np.uint64) * np.uint64(2654435761)).astype(np.uint32)
    np.testing.assert_array_equal(h, ((ux ^ uy) & np.uint32(2 ** 19 - 1)).astype(np.int64))


@pytest.mark.parametrize("name", ALL_CASES)
def test_forward_matches_reference(name):
    g = load(name)
    fwd = O.gngf_forward(params_of(g), g["x"], oracle_cfg(g))
    np.testing.assert_array_equal(fwd["idx"], g["idx"])
    assert rel_err(np.transpose(fwd["feat"], (0, 1, 2, 3)), g["feat"]) < FWD_TOL
    assert rel_err(fwd["enc"], g["enc"]) < FWD_TOL
    assert rel_err(fwd["rgb"], g["rgb"]) < FWD_TOL
    if not g["cfg"]["use_hash"]:
        assert rel_err(fwd["topv"], g["topv"]) < FWD_TOL
        assert rel_err(fwd["probs"][:4], g["probs_head"]) < FWD_TOL
        assert rel_err(fwd["ret_probs"][:4], g["ret_probs_head"]) < FWD_TOL
        assert rel_err(fwd["pbar"], g["pbar"]) < FWD_TOL


@pytest.mark.parametrize("name", GNGF_CASES)
def test_loss_matches_reference(name):
    g = load(name)
    lc = loss_cfg(g)
    fwd = O.gngf_forward(params_of(g), g["x"], oracle_cfg(g))
    coll = g["coll_losses"] if "coll_losses" in g else None
    total, mse, level, _ = O.total_loss(fwd["rgb"], g["y"], fwd["pbar"], lc["gamma"], lc["epsilon"],
                                        lc["l_mse"], lc["l_js_kl"], lc["l_collisions"], coll)
    assert abs(mse - g["mse"]) < FWD_TOL * abs(g["mse"])
    assert rel_err(level, g["kl_levels"]) < 1e-4     # ln() of ~1/N means: a few ulps of cancellation
    assert abs(total - g["loss"]) < 1e-5 * abs(g["loss"])


@pytest.mark.parametrize("name", ALL_CASES)
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_backward_matches_reference(name, dtype):
    g = load(name)
    p = params_of(g, dtype)
    x, y = g["x"].astype(dtype), g["y"].astype(dtype)
    cfg = oracle_cfg(g)
    fwd = O.gngf_forward(p, x, cfg)
    if dtype == np.float64 and not g["cfg"]["use_hash"]:
        # fp64 logits may order a near-tie differently; the closed form must use the reference's selection
        if not np.array_equal(fwd["idx"], g["idx"]):
            pytest.skip("fp64 recompute picked a different top-k on a near-tie")
    grads = O.gngf_backward(p, x, y, cfg, fwd, loss_cfg(g))
    L = g["cfg"]["L"]
    for l in range(L):
        assert rel_err(grads["tables"][l], g[f"grad.encoding._hash_tables.{l}.weight"]) < GRAD_TOL, l
    for i in range(len(p["mlp_w"])):
        assert rel_err(grads["mlp_w"][i], g[f"grad.mlp.{i}.0.weight"]) < GRAD_TOL, i
        assert rel_err(grads["mlp_b"][i], g[f"grad.mlp.{i}.0.bias"]) < GRAD_TOL, i
    if not g["cfg"]["use_hash"]:
        for i in range(len(p["hpd_w"])):
            assert rel_err(grads["hpd_w"][i], g[f"grad.HPD.module_list.{i}.0.weight"]) < GRAD_TOL, i
            assert rel_err(grads["hpd_b"][i], g[f"grad.HPD.module_list.{i}.0.bias"]) < GRAD_TOL, i


@pytest.mark.parametrize("name", ALL_CASES)
def test_calc_hash_collisions_matches_reference(name):
    g = load(name)
    if g["cfg"]["use_hash"]:
        coll, minp = O.calc_hash_collisions_hash_mode(g["idx"], g["n_ls"], g["cfg"]["T"])
    else:
        coll, minp = O.calc_hash_collisions(g["idx"].astype(np.float32), g["n_ls"], g["cfg"]["T"])
    np.testing.assert_allclose(coll, g["chc_collisions"], rtol=0, atol=0)
    np.testing.assert_array_equal(minp, g["chc_min_possible"])


@pytest.mark.parametrize("name", COUNTS_CASES)
def test_calc_counts_per_level_matches_reference(name):
    """f-4: _calc_counts_per_level (models.py:530-566) incl. its (point index -> flattened (p v) slot) indexing."""
    g = load(name)
    _, grid = O.scale_to_grid(g["x"], g["n_ls"])
    hashed = g["idx"] if g["cfg"]["use_hash"] else g["idx"][..., 0]
    got = O.calc_counts_per_level(hashed, grid)
    assert got == golden_counts(g)


@pytest.mark.parametrize("name", ["cfg2_small", "cfg2_topk_only", "l16_t1024", "k20"])
def test_torch_port_matches_reference(name):
    """oracle/torch_port.py (the timed CPU baseline) reproduces the reference's forward, loss and gradients."""
    import torch
    import torch_port as TP
    g = load(name)
    c = g["cfg"]
    p = params_of(g)
    tp = {k: [torch.from_numpy(a.copy()).requires_grad_() for a in v] for k, v in p.items()}
    cfg = {"n_ls": g["n_ls"], "topk_k": c["K"], "topk_only": c["topk_only"]}
    rgb, probs, idx = TP.forward(tp, torch.from_numpy(g["x"]), cfg)
    total, mse, levels = TP.loss(rgb, torch.from_numpy(g["y"]), probs, c["gamma"], c["epsilon"], c["l_mse"], c["l_js_kl"])
    total.backward()
    assert np.array_equal(idx.numpy(), g["idx"])
    assert rel_err(rgb.detach().numpy(), g["rgb"]) < FWD_TOL
    assert abs(float(total) - g["loss"]) < 1e-5 * abs(g["loss"])
    for l in range(c["L"]):
        assert rel_err(tp["tables"][l].grad.numpy(), g[f"grad.encoding._hash_tables.{l}.weight"]) < GRAD_TOL
    for i in range(len(tp["hpd_w"])):
        assert rel_err(tp["hpd_w"][i].grad.numpy(), g[f"grad.HPD.module_list.{i}.0.weight"]) < GRAD_TOL
    for i in range(len(tp["mlp_w"])):
        assert rel_err(tp["mlp_w"][i].grad.numpy(), g[f"grad.mlp.{i}.0.weight"]) < GRAD_TOL
