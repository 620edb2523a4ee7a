"""GPU: the whole hot path (forward + loss + backward through the drop-in module) against golden vectors of
the unmodified reference and against the oracle.  Tolerances (BASELINE.json north_star): top-k indices exact,
forward 1e-5 relative, gradients 1e-4 relative."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import gngf_oracle as O  # noqa: E402
from golden_util import (ALL_CASES, COUNTS_CASES, GNGF_CASES, golden_counts, load, loss_cfg, oracle_cfg, params_of,  # noqa: E402
                         rel_err)
from parity_util import decoder_masks, build_net, reset_flags, run_step  # noqa: E402

pytestmark = pytest.mark.gpu
FWD_TOL, GRAD_TOL = 1e-5, 1e-4


@pytest.fixture(autouse=True)
def _flags():
    yield
    reset_flags()


def _check_grads(out, g, tol=GRAD_TOL):
    for k in [k for k in g if k.startswith("grad.")]:
        name = k[len("grad."):]
        assert name in out["grads"], name
        assert rel_err(out["grads"][name], g[k]) < tol, (name, rel_err(out["grads"][name], g[k]))


@pytest.mark.parametrize("name", ALL_CASES)
def test_step_matches_reference_golden(name):
    g = load(name)
    net = build_net(g)
    out = run_step(net, g)
    assert np.array_equal(out["idx"], g["idx"])
    assert out["idx"].dtype == np.int64
    assert rel_err(out["rgb"], g["rgb"]) < FWD_TOL
    st = out["state"]
    assert rel_err(st.mlp_acts[0].cpu().numpy(), g["enc"]) < FWD_TOL
    if not g["cfg"]["use_hash"]:
        assert int(st.err_flag.item()) == 0
        assert rel_err(out["pbar"], g["pbar"]) < FWD_TOL
        assert abs(out["mse"] - g["mse"]) < FWD_TOL * abs(g["mse"])
        assert rel_err(out["kl_levels"], g["kl_levels"]) < 1e-4
    assert abs(out["loss"] - g["loss"]) < 1e-5 * abs(g["loss"])
    _check_grads(out, g)


@pytest.mark.parametrize("name", ["cfg2_small", "cfg2_topk_only", "k20"])
def test_materialised_probs_match_reference_and_lazy_path(name):
    g = load(name)
    net = build_net(g)
    rgb, probs, idx, _ = net(torch.from_numpy(g["x"]).cuda(), 1.0)
    dense = probs.materialize()
    assert tuple(dense.shape) == tuple(probs.shape) == (g["x"].shape[0], g["cfg"]["L"], 4, g["ret_probs_head"].shape[-1])
    assert rel_err(dense[:4].detach().cpu().numpy(), g["ret_probs_head"]) < FWD_TOL
    # the loss through the materialised tensor gives the same gradients as through the symbolic column sums
    out_dense = run_step(net, g, materialize=True)
    _check_grads(out_dense, g)


def test_full_size_cfg2_against_oracle():
    """Config 2 at full batch size (P = 57 404 strawberry pixels, ID 4061 parameters)."""
    g = load("cfg2_small")
    img = np.load(os.path.join(os.path.dirname(__file__), "golden", "strawberry_u8.npz"))["img"]
    h, w = img.shape[:2]
    X = np.stack(np.meshgrid(range(h), range(w), indexing="ij"), -1).reshape(-1, 2)
    x_all = (torch.tensor(X).float() / (max(w, h) - 1)).numpy()
    y_all = (torch.tensor(img.reshape(-1, 3) / 255).float()).numpy()
    perm = np.random.default_rng(65535).permutation(x_all.shape[0])[:57404]
    g = dict(g)
    g["x"], g["y"] = x_all[perm], y_all[perm]
    net = build_net(g)
    out = run_step(net, g)
    assert out["state"].lat.num_nodes == 34 * 23
    p = params_of(g)
    fwd = O.gngf_forward(p, g["x"], oracle_cfg(g))
    assert np.array_equal(out["idx"], fwd["idx"])
    assert rel_err(out["rgb"], fwd["rgb"]) < FWD_TOL
    assert rel_err(out["pbar"], fwd["pbar"]) < FWD_TOL
    # gradients: at 918 464 rows the float32 oracle's own row sums are only good to ~1e-3, so the reference
    # values come from the float64 oracle differentiating the same top-k selection
    p64 = params_of(g, np.float64)
    cfg64 = dict(oracle_cfg(g), force_idx=fwd["idx"])
    x64, y64 = g["x"].astype(np.float64), g["y"].astype(np.float64)
    fwd64 = O.gngf_forward(p64, x64, cfg64)
    # ... and the same ReLU pattern of the decoder: of the 7.3 M hidden pre-activations a handful lie within the
    # forward's rounding error (1e-6) of zero, where the derivative is discontinuous
    masks = decoder_masks(out["state"])
    if masks is not None:
        for i in (1, 2):
            diff = masks[i] != (fwd64["mlp_acts"][i] > 0)
            assert diff.mean() < 1e-5, diff.sum()
        cfg64["force_mlp_masks"] = masks
    grads = O.gngf_backward(p64, x64, y64, cfg64, fwd64, loss_cfg(g))
    for l in range(4):
        assert rel_err(out["grads"][f"encoding._hash_tables.{l}.weight"], grads["tables"][l]) < GRAD_TOL
    for i in range(4):
        assert rel_err(out["grads"][f"HPD.module_list.{i}.0.weight"], grads["hpd_w"][i]) < GRAD_TOL
        assert rel_err(out["grads"][f"HPD.module_list.{i}.0.bias"], grads["hpd_b"][i]) < GRAD_TOL
    for i in range(3):
        assert rel_err(out["grads"][f"mlp.{i}.0.weight"], grads["mlp_w"][i]) < GRAD_TOL
        assert rel_err(out["grads"][f"mlp.{i}.0.bias"], grads["mlp_b"][i]) < GRAD_TOL


@pytest.mark.parametrize("name", COUNTS_CASES)
def test_counts_per_level_match_reference(name):
    """f-4: forward(..., should_calc_counts=True) returns the reference's per-level histograms (models.py:431-439,
    530-566) -- GPU de-duplication (k8_collisions.cu) instead of np.unique(axis=0) on the host."""
    g = load(name)
    net = build_net(g)
    x = torch.from_numpy(g["x"]).cuda()
    n0 = __import__("collision_handling_in_instantngp_b200").launch_count()
    _, _, idx, counts = net(x, 1.0, should_calc_counts=True)
    assert counts == golden_counts(g)
    assert np.array_equal(idx.cpu().numpy(), g["idx"])
    # ... and through the stand-alone method on materialised tensors, as a caller of the reference's API would
    scaled, grid = net._scale_to_grid(x)
    hashed = idx if g["cfg"]["use_hash"] else idx[..., 0]
    assert net._calc_counts_per_level(hashed, grid) == golden_counts(g)
    # a non-contiguous / odd-stride view takes the same route
    assert net._calc_counts_per_level(hashed.clone(), grid) == golden_counts(g)
    # inputs the kernel cannot place (corners outside every box) fall back to the reference's host route
    far = grid + 1.0e6
    assert sum(sum(d.values()) for d in net._calc_counts_per_level(hashed, far)) > 0


def test_state_dict_keys_and_optimizer_groups():
    g = load("cfg2_small")
    net = build_net(g)
    keys = list(net.state_dict().keys())
    assert keys == [k[len("param."):] if k.startswith("param.") else k for k in keys]
    expect = ["_batch_norm.weight", "_batch_norm.bias", "_batch_norm.running_mean", "_batch_norm.running_var",
              "_batch_norm.num_batches_tracked"] + \
             [f"HPD.module_list.{i}.0.{n}" for i in range(4) for n in ("weight", "bias")] + \
             [f"encoding._hash_tables.{l}.weight" for l in range(4)] + \
             [f"mlp.{i}.0.{n}" for i in range(3) for n in ("weight", "bias")]
    assert keys == expect
    # functions.py:96-127
    opt = torch.optim.Adam([{"params": net.encoding.parameters(), "lr": 1e-4},
                            {"params": net.HPD.parameters(), "lr": 1e-3},
                            {"params": net.mlp.parameters(), "lr": 1e-3}], betas=(0.9, 0.99), eps=1e-15)
    out = run_step(net, g)
    opt.step()
    out2 = run_step(net, g)
    assert out2["loss"] != out["loss"]


def test_frozen_hpd_and_empty_batch():
    g = load("cfg2_small")
    net = build_net(g)
    for p in net.HPD.parameters():
        p.requires_grad = False
    out = run_step(net, g)
    assert not any(k.startswith("HPD") for k in out["grads"])
    _check_grads({"grads": {**out["grads"], **{k[5:]: g[k] for k in g if k.startswith("grad.HPD")}}}, g)


def test_out_of_bounds_coordinates_are_flagged():
    g = load("cfg2_small")
    net = build_net(g)
    net.set_coord_bounds((0.0, 0.0), (0.5, 0.5))
    x = torch.from_numpy(g["x"]).cuda()
    net(x, 1.0)
    assert int(net.last_state.err_flag.item()) == 1
    net.set_coord_bounds(None)
    net(x, 1.0)
    assert int(net.last_state.err_flag.item()) == 0


@pytest.mark.parametrize("name", ["cfg2_topk_only", "l8_t4096_topk_only"])
def test_streaming_tensor_core_path_matches_reference_golden(name):
    """should_keep_topk_only=True through the tcgen05 streaming HPD (logits never materialised in the forward,
    recomputed chunk-wise in the backward) against the reference's golden step."""
    from collision_handling_in_instantngp_b200 import ops
    g = load(name)
    net = build_net(g)
    ops.FORCE_STREAMING = True
    try:
        out = run_step(net, g)
    finally:
        ops.FORCE_STREAMING = None
    st = out["state"]
    assert st.uprobs is None and st.row_max is not None
    assert np.array_equal(out["idx"], g["idx"])
    assert rel_err(out["rgb"], g["rgb"]) < FWD_TOL
    assert rel_err(out["pbar"], g["pbar"]) < FWD_TOL
    assert abs(out["loss"] - g["loss"]) < 1e-5 * abs(g["loss"])
    _check_grads(out, g)


@pytest.mark.parametrize("name", ["cfg2_topk_only", "l8_t4096_topk_only"])
def test_active_node_evaluation_matches_reference_golden(name):
    """The streaming HPD evaluated only on the lattice nodes the batch touches (k11_active_nodes.cu; what the 8192^2
    lattice of BASELINE.json configs[3] runs) against the reference's golden step and against the evaluation of the
    whole box."""
    from collision_handling_in_instantngp_b200 import ops
    g = load(name)
    net = build_net(g)
    ops.FORCE_STREAMING = True
    try:
        out_box = run_step(net, g)
        ops.FORCE_ACTIVE_NODES = True
        out = run_step(net, g)
    finally:
        ops.FORCE_STREAMING = None
        ops.FORCE_ACTIVE_NODES = None
    st = out["state"]
    assert out_box["state"].node_ids is None
    assert st.node_ids is not None and 0 < st.node_ids.shape[0] < st.lat.num_nodes
    assert st.hpd_acts[0].shape[0] == st.node_ids.shape[0] and st.utopv.shape[0] == st.lat.num_nodes
    assert np.array_equal(out["idx"], g["idx"])
    assert rel_err(out["rgb"], g["rgb"]) < FWD_TOL
    assert rel_err(out["pbar"], g["pbar"]) < FWD_TOL
    assert abs(out["loss"] - g["loss"]) < 1e-5 * abs(g["loss"])
    _check_grads(out, g)
    assert np.array_equal(out["idx"], out_box["idx"])
    assert rel_err(out["rgb"], out_box["rgb"]) < 2e-6             # (the split of T over CTAs follows the row count)
    for k in out["grads"]:
        assert rel_err(out["grads"][k], out_box["grads"][k]) < GRAD_TOL, k   # (two-plane products: ~1e-5 each)


@pytest.mark.parametrize("regime", ["unit_scale_inputs", "raw_init"])
def test_active_nodes_on_a_large_lattice_equal_the_whole_box(regime):
    """Active-node evaluation at scale: the 8192-resolution lattice of BASELINE.json configs[3] (16 levels), a quarter
    box (4098^2 = 16.8 M nodes, so that the whole-box evaluation it is compared with fits comfortably), 2^20 pixel-lattice
    points: the touched-node list, the scatter back into (U, K) and the node-list backward against the evaluation of
    every node of the box.

    The HPD is fed integer lattice coordinates (models.py:416-418) -- up to 4 097 here -- so with nn.Linear's initial
    weights its logits are O(1e4) and the softmax is one-hot ("raw_init"); "unit_scale_inputs" scales the first layer by
    1/4096 -- logits O(1), a flat softmax over all slots.  In both regimes the forward must agree exactly and the table /
    decoder gradients to 1e-4.  The HPD gradients are sums over 7.5 M (active) or 16.8 M (whole box) rows accumulated in
    fp32 in a different grouping by the two evaluations: they agree to the rounding noise of such a sum (measured 7e-5 and
    2.4e-4 -- the flat regime adds 1 000 comparable terms per row), bounded here at 5e-4; the accuracy of the streaming
    backward itself is pinned against float64 in tests/test_kernels_gpu.py."""
    from collision_handling_in_instantngp_b200 import ops
    from collision_handling_in_instantngp_b200.loss import fused_total_loss
    from collision_handling_in_instantngp_b200.models import GeneralNeuralGaugeFields
    torch.manual_seed(65535)
    T, K = 1024, 4
    net = GeneralNeuralGaugeFields(input_dim=2, hash_table_size=T, num_levels=16, n_min=16, n_max=8192,
                                   MLP_hidden_layers_widths=[64, 64], HPD_hidden_layers_widths=[32, 64, 128],
                                   HPD_out_features=T, topk_k=K, should_keep_topk_only=True)
    with torch.no_grad():
        for t in net.encoding.tables():
            t.mul_(300.0)                          # table gradients well above fp32 noise (as in the goldens)
        if regime == "unit_scale_inputs":
            net.HPD.module_list[0][0].weight.mul_(1.0 / 4096)
    half = 4096
    net.set_coord_bounds((0.0, 0.0), (half / 8191, half / 8191))
    rng = np.random.default_rng(3)
    flat = rng.permutation((half + 1) * (half + 1))[: 2 ** 20]
    x = torch.from_numpy(np.stack([flat // (half + 1), flat % (half + 1)], 1).astype(np.float32) / np.float32(8191)).cuda()
    y = torch.from_numpy(rng.random((2 ** 20, 3), dtype=np.float32)).cuda()

    def run(active):
        ops.FORCE_ACTIVE_NODES = active
        try:
            net.zero_grad()
            rgb, probs, idx, _ = net(x, 1.0)
            loss, _, _ = fused_total_loss(rgb, y, probs.colsum, 4 * x.shape[0], -2.0, 1.0, 1.0, 1.0)
            loss.backward()
            st = net.last_state
            n_ids = None if st.node_ids is None else int(st.node_ids.shape[0])
            return (rgb.detach().clone(), idx.clone(), probs.colsum.detach().clone(), float(loss),
                    {k: v.grad.detach().clone() for k, v in net.named_parameters() if v.grad is not None}, n_ids,
                    st.lat.num_nodes)
        finally:
            ops.FORCE_ACTIVE_NODES = None

    rgb_a, idx_a, cs_a, loss_a, g_a, n_a, U = run(True)
    rgb_b, idx_b, cs_b, loss_b, g_b, n_b, _ = run(False)
    assert U >= half * half
    assert n_b is None and 0 < n_a < 0.75 * U
    errs = {k: float((g_a[k] - g_b[k]).abs().max() / (g_b[k].abs().max() + 1e-30)) for k in g_b}
    hpd_worst = max(v for k, v in errs.items() if k.startswith("HPD"))
    rest_worst = max(v for k, v in errs.items() if not k.startswith("HPD"))
    print(f"\n[{regime}] lattice nodes {U}, touched {n_a} ({n_a / U:.1%}); gradient differences active vs whole box: HPD "
          f"{hpd_worst:.2e}, tables + decoder {rest_worst:.2e}; HPD per parameter "
          f"{({k[len('HPD.module_list.'):]: float(f'{v:.1e}') for k, v in errs.items() if k.startswith('HPD')})}")
    assert torch.equal(idx_a, idx_b)
    assert float((rgb_a - rgb_b).abs().max()) < 2e-6
    assert float(((cs_a - cs_b).abs() / cs_b.abs()).max()) < 1e-5
    assert abs(loss_a - loss_b) < 1e-5 * abs(loss_b)
    assert rest_worst < GRAD_TOL, errs
    assert hpd_worst < 5e-4, errs


@pytest.mark.parametrize("mix", [False, None])
def test_active_node_evaluation_with_every_mix_mode(mix):
    """Untouched nodes carry placeholder selections; no mix mode may turn them into NaN gradients."""
    from collision_handling_in_instantngp_b200 import ops
    g = load("cfg2_topk_only")
    net = build_net(g, mix_mode=mix)
    ops.FORCE_STREAMING = True
    try:
        out_box = run_step(net, g)
        ops.FORCE_ACTIVE_NODES = True
        out = run_step(net, g)
    finally:
        ops.FORCE_STREAMING = None
        ops.FORCE_ACTIVE_NODES = None
    assert out["state"].node_ids is not None
    assert np.array_equal(out["idx"], out_box["idx"]) and rel_err(out["rgb"], out_box["rgb"]) < 2e-6
    for k in out["grads"]:
        assert np.isfinite(out["grads"][k]).all(), k
        assert rel_err(out["grads"][k], out_box["grads"][k]) < GRAD_TOL, k


@pytest.mark.parametrize("name", ["cfg2_topk_only", "l8_t4096_topk_only"])
def test_fused_streaming_backward_equals_chunked_recompute(name):
    """k2_hpd_tc_bwd.cu (dlogits never materialised) against the chunked recompute through the plain GEMM."""
    from collision_handling_in_instantngp_b200 import ops
    g = load(name)
    net = build_net(g)
    ops.FORCE_STREAMING = True
    try:
        out_f = run_step(net, g)
        ops.STREAM_BWD_FUSED = False
        out_c = run_step(net, g)
    finally:
        ops.FORCE_STREAMING = None
        ops.STREAM_BWD_FUSED = True
    for k in out_c["grads"]:
        assert rel_err(out_f["grads"][k], out_c["grads"][k]) < 1e-4, k


@pytest.mark.parametrize("name", ["cfg2_small", "cfg2_topk_only", "cfg2_epoch1", "js_only", "kl_only", "l16_t1024"])
def test_fused_loss_kernel_matches_reference_loss(name):
    """k7_loss.cu (value + adjoints) against the reference's Loss / autograd recorded in the goldens."""
    from collision_handling_in_instantngp_b200.loss import fused_total_loss, total_loss
    g = load(name)
    c = g["cfg"]
    net = build_net(g)
    x, y = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["y"]).cuda()
    coll = None
    if "coll_losses" in g:
        coll = c["l_collisions"] * torch.from_numpy(g["coll_losses"]).cuda()
    net.zero_grad()
    rgb, probs, _, _ = net(x, 1.0)
    rows = probs.shape[0] * probs.shape[2]
    total, mse, levels = fused_total_loss(rgb, y, probs.colsum, rows, c["gamma"], c["epsilon"], c["l_mse"], c["l_js_kl"],
                                          coll)
    assert abs(float(total) - g["loss"]) < 1e-5 * abs(g["loss"])
    assert abs(float(mse) - g["mse"]) < 1e-5 * abs(g["mse"])
    assert rel_err(levels.cpu().numpy(), g["kl_levels"]) < 1e-4
    total.backward()
    out = {"grads": {k: v.grad.detach().cpu().numpy() for k, v in net.named_parameters() if v.grad is not None}}
    _check_grads(out, g)
    # and against the torch restatement of the same formula
    t2, _, _ = total_loss(rgb.detach(), y, probs.colsum.detach(), rows, c["gamma"], c["epsilon"], c["l_mse"],
                          c["l_js_kl"], coll)
    assert abs(float(t2) - float(total)) < 1e-5 * abs(float(t2))


@pytest.mark.parametrize("name", ALL_CASES)
def test_calc_hash_collisions_matches_reference(name):
    g = load(name)
    net = build_net(g)
    idx = torch.from_numpy(g["idx"]).cuda()
    coll, minp = net.calc_hash_collisions(idx if g["cfg"]["use_hash"] else idx.float())
    # counts are exact integers; the mean over the K columns is a float32 reduction (1 ulp of order freedom)
    np.testing.assert_allclose(coll.cpu().numpy(), g["chc_collisions"], rtol=1e-6, atol=0)
    assert np.array_equal(minp.cpu().numpy(), g["chc_min_possible"])
    if not g["cfg"]["use_hash"]:
        # int64 input and a buffer with garbage (what torch.empty may hold) agree with exact counting
        coll_i, _ = net.calc_hash_collisions(idx)
        np.testing.assert_allclose(coll_i.cpu().numpy(), g["chc_collisions"], rtol=1e-6, atol=0)
        junk = idx.float().clone()
        junk[::3, :, :, 0] = 1e30
        junk[1::7, :, :, -1] = -0.5
        coll_j, _ = net.calc_hash_collisions(junk)
        ref, _ = O.calc_hash_collisions(junk.cpu().numpy(), g["n_ls"], g["cfg"]["T"])
        np.testing.assert_allclose(coll_j.cpu().numpy(), ref, rtol=1e-6, atol=0)


@pytest.mark.parametrize("name", ["cfg2_small", "cfg2_topk_only", "k20", "mix_raw"])
def test_fused_small_lattice_hpd_equals_general_path(name):
    """k2_hpd_small.cu (one CTA per 8 nodes, whole HPD forward / backward) against the layer-by-layer path."""
    from collision_handling_in_instantngp_b200 import ops
    g = load(name)
    net = build_net(g)
    out_small = run_step(net, g)
    assert out_small["state"].hpd_small
    saved = ops.SMALL_LATTICE_MAX_NODES
    ops.SMALL_LATTICE_MAX_NODES = 0
    try:
        out_gen = run_step(net, g)
    finally:
        ops.SMALL_LATTICE_MAX_NODES = saved
    assert not out_gen["state"].hpd_small
    assert np.array_equal(out_small["idx"], out_gen["idx"])
    assert rel_err(out_small["rgb"], out_gen["rgb"]) < 1e-6
    for k in out_gen["grads"]:
        assert rel_err(out_small["grads"][k], out_gen["grads"][k]) < 2e-5, k


@pytest.mark.parametrize("name", ["cfg2_small", "cfg2_topk_only", "cfg2_epoch1", "k1", "k20", "mix_raw",
                                  "mix_weighted_avg", "bw_leaky"])
def test_node_passes_folded_into_the_small_hpd_kernels(name):
    """k2_hpd_small.cu with the encoding's per-level-node passes folded in (gngf_hpd_small_fwd_enc / _bwd_enc) against
    the separate node_features_fwd / _bwd launches: same selections, same features, same gradients."""
    from collision_handling_in_instantngp_b200 import ops
    g = load(name)
    net = build_net(g)
    out_fused = run_step(net, g)
    assert out_fused["state"].hpd_small and ops.SMALL_FUSE_NODE_PASSES
    ops.SMALL_FUSE_NODE_PASSES = False
    try:
        out_sep = run_step(net, g)
    finally:
        ops.SMALL_FUSE_NODE_PASSES = True
    assert np.array_equal(out_fused["idx"], out_sep["idx"])
    assert np.array_equal(out_fused["rgb"], out_sep["rgb"])       # identical arithmetic per level node
    for k in out_sep["grads"]:
        assert rel_err(out_fused["grads"][k], out_sep["grads"][k]) < 2e-6, k
    _check_grads(out_fused, g)


def test_graphed_trainer_matches_eager_steps():
    """trainer.GraphedTrainer (whole step in CUDA graphs, fused Adam, loss adjoints seeding the backward, double-buffered
    pipelined inputs) against the same steps driven eagerly through the module API with torch.optim.Adam."""
    from collision_handling_in_instantngp_b200.loss import total_loss
    from collision_handling_in_instantngp_b200.optim import FusedAdam
    from collision_handling_in_instantngp_b200.trainer import GraphedTrainer
    g = load("cfg2_small")
    c = g["cfg"]
    P = g["x"].shape[0]
    rng = np.random.default_rng(5)
    batches = []
    for _ in range(6):
        perm = rng.permutation(P)
        batches.append((torch.from_numpy(g["x"][perm]).pin_memory(), torch.from_numpy(g["y"][perm]).pin_memory()))

    def groups(net, cls, **kw):
        return cls([{"params": net.encoding.parameters(), "lr": 1e-3}, {"params": net.HPD.parameters(), "lr": 1e-3,
                    "weight_decay": 1e-6}, {"params": net.mlp.parameters(), "lr": 1e-3, "weight_decay": 1e-6}],
                   betas=(0.9, 0.99), eps=1e-15, **kw)

    # eager reference
    net_e = build_net(g)
    net_e.set_coord_bounds((0.0, 0.0), (1.0, 1.0))
    opt_e = groups(net_e, torch.optim.Adam)
    losses_e = []
    for x, y in batches:
        opt_e.zero_grad()
        rgb, probs, _, _ = net_e(x.cuda(), 1.0)
        loss, _, _ = total_loss(rgb, y.cuda(), probs.colsum, 4 * P, c["gamma"], c["epsilon"], c["l_mse"], c["l_js_kl"])
        loss.backward()
        opt_e.step()
        losses_e.append(float(loss))

    for mode in ("blocking", "pipelined"):
        net_g = build_net(g)
        net_g.set_coord_bounds((0.0, 0.0), (1.0, 1.0))
        opt_g = groups(net_g, FusedAdam)
        init = {k: v.detach().clone() for k, v in net_g.state_dict().items()}
        # warm-up on random coordinates / zero targets (no sample batch): the constructor must leave the parameters, the
        # buffers and the optimizer state (moments, step counters) exactly as it found them
        tr = GraphedTrainer(net_g, opt_g, points=P, gamma=c["gamma"], epsilon=c["epsilon"], l_mse=c["l_mse"],
                            l_js_kl=c["l_js_kl"], warmup_steps=2)
        for k, v in net_g.state_dict().items():
            assert torch.equal(v, init[k]), k
        for st in opt_g.state.values():
            assert float(st["step"]) == 0.0 and not st["exp_avg"].any() and not st["exp_avg_sq"].any()
        losses_g = []
        if mode == "blocking":
            for x, y in batches:
                losses_g.append(float(tr.step(x, y)))
        else:
            for x, y in batches:
                prev = tr.step_pipelined(x, y)
                if prev is not None:
                    losses_g.append(prev)
            losses_g.append(tr.flush())
        np.testing.assert_allclose(losses_g, losses_e, rtol=2e-5)
        for (k, a), (_, b) in zip(net_g.state_dict().items(), net_e.state_dict().items()):
            if a.dtype.is_floating_point:
                assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 2e-4, (mode, k)


def test_graphed_trainer_raises_on_out_of_bounds_coordinates():
    """A batch that violates the bounds promised to set_coord_bounds is clamped by the kernels; the sticky device flag
    must surface as GngfError from the trainer (and from net.check_errors() in eager use)."""
    from collision_handling_in_instantngp_b200 import GngfError
    from collision_handling_in_instantngp_b200.optim import FusedAdam
    from collision_handling_in_instantngp_b200.trainer import GraphedTrainer
    g = load("cfg2_small")
    c = g["cfg"]
    P = g["x"].shape[0]
    net = build_net(g)
    net.set_coord_bounds((0.0, 0.0), (0.5, 0.5))
    x, y = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["y"]).cuda()
    net(x, 1.0)
    with pytest.raises(GngfError):
        net.check_errors()
    net.check_errors()                       # cleared by the raise
    opt = FusedAdam(net.parameters(), lr=1e-3)
    tr = GraphedTrainer(net, opt, points=P, gamma=c["gamma"], epsilon=c["epsilon"], warmup_steps=1)
    tr.check_errors()                        # the in-bounds warm-up batches left no flag behind
    tr.step(x, y)
    with pytest.raises(GngfError):
        tr.check_errors()
    net._err_flag_sticky.zero_()
    tr.check_every = 1
    tr.step_pipelined(x.cpu().pin_memory(), y.cpu().pin_memory())
    with pytest.raises(GngfError):
        tr.flush()


def test_side_stream_forks_do_not_change_results():
    """ops.CONCURRENT (independent kernels on side streams) against the fully serial schedule: identical outputs and
    gradients (the kernels and their inputs are the same; only the launch order across streams differs)."""
    from collision_handling_in_instantngp_b200 import ops
    g = load("cfg2_small")
    net = build_net(g)
    out_c = run_step(net, g)
    ops.CONCURRENT = False
    try:
        out_s = run_step(net, g)
    finally:
        ops.CONCURRENT = True
    assert np.array_equal(out_c["idx"], out_s["idx"])
    assert np.array_equal(out_c["rgb"], out_s["rgb"])
    for k in out_s["grads"]:
        # (float atomics inside single kernels may reorder between runs: compare to rounding level)
        assert rel_err(out_c["grads"][k], out_s["grads"][k]) < 1e-5, k
