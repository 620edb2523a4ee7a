"""GPU: per-kernel known-answer tests through the C ABI, against oracle/gngf_oracle.py and the goldens.
Bar: integer corners, hashes and top-k indices bit-exact; fp32 values within 1e-5 relative."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import gngf_oracle as O  # noqa: E402
from golden_util import ALL_CASES, load, rel_err  # noqa: E402

from collision_handling_in_instantngp_b200 import ops  # noqa: E402
from collision_handling_in_instantngp_b200.lattice import build_lattice, level_resolutions  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _edge_coords(n=4096, seed=0):
    rng = np.random.default_rng(seed)
    x = rng.random((n, 2), dtype=np.float32)
    x[:8] = [[0, 0], [1, 1], [1, 0], [0, 1], [0.5, 0.5], [1 / 3, 2 / 3], [np.nextafter(np.float32(1), np.float32(0))] * 2,
             [338 / 507, 1.0]]
    # exact lattice hits of an 8192 grid: i / 8191
    x[8:1032, 0] = (rng.integers(0, 8192, 1024) / np.float32(8191)).astype(np.float32)
    x[8:1032, 1] = (rng.integers(0, 8192, 1024) / np.float32(8191)).astype(np.float32)
    return x


@pytest.mark.parametrize("levels", [(8, 32, 4), (16, 508, 16), (16, 8192, 16), (2, 1024, 10)])
def test_corners_bit_exact(levels):
    n_ls = level_resolutions(*levels)
    assert np.array_equal(n_ls, O.level_resolutions(*levels))
    x = _edge_coords()
    scaled, grid = ops.corners_fwd(torch.from_numpy(x).to(DEV), build_lattice(n_ls))
    s_ref, g_ref = O.scale_to_grid(x, n_ls)
    assert np.array_equal(scaled.cpu().numpy(), s_ref)
    assert np.array_equal(grid.cpu().numpy(), g_ref)


@pytest.mark.parametrize("name", ALL_CASES)
def test_corners_match_reference_goldens(name):
    g = load(name)
    scaled, grid = ops.corners_fwd(torch.from_numpy(g["x"]).to(DEV), build_lattice(g["n_ls"]))
    assert np.array_equal(scaled.cpu().numpy(), g["scaled"])
    assert np.array_equal(grid.cpu().numpy(), g["grid"])


@pytest.mark.parametrize("levels,P", [((8, 32, 4), 57), ((16, 508, 16), 4096), ((16, 8192, 16), 4096), ((2, 1024, 10), 1)])
def test_active_nodes_are_the_distinct_corner_nodes(levels, P):
    """mark + compact (k11_active_nodes.cu): ascending ids of exactly the lattice nodes the oracle's corners touch."""
    n_ls = level_resolutions(*levels)
    x = _edge_coords()[:P]
    lat = build_lattice(n_ls)
    ids = ops.active_nodes(torch.from_numpy(x).to(DEV), lat)
    _, g_ref = O.scale_to_grid(x, n_ls)                              # (P,2,L,4) integer-valued fp32
    u = (g_ref[:, 0].astype(np.int64) - lat.ox) * lat.wy + (g_ref[:, 1].astype(np.int64) - lat.oy)
    assert ids.dtype == torch.int32
    assert np.array_equal(ids.cpu().numpy().astype(np.int64), np.unique(u))


def test_scatter_node_rows_and_first_layer_on_node_lists():
    lat = build_lattice(level_resolutions(16, 508, 16))
    U = lat.num_nodes
    g = torch.Generator(device=DEV).manual_seed(3)
    ids = torch.sort(torch.randperm(U, device=DEV, generator=g)[:1000])[0].int()
    src = torch.randn((1000, 4), device=DEV, generator=g)
    dst = torch.ones((U, 4), device=DEV)
    ops.scatter_node_rows(ids, src, dst)
    ref = torch.ones((U, 4), device=DEV)
    ref[ids.long()] = src
    assert torch.equal(dst, ref)
    isrc = torch.randint(0, 1 << 19, (1000, 4), device=DEV, generator=g, dtype=torch.int32)
    idst = torch.zeros((U, 4), device=DEV, dtype=torch.int32)
    ops.scatter_node_rows(ids, isrc, idst)
    assert torch.equal(idst[ids.long()], isrc) and int((idst != 0).sum()) == int((isrc != 0).sum())
    # first HPD layer on the list == the rows of the full evaluation; its backward == the full one on scattered adjoints
    w0 = torch.randn((32, 2), device=DEV, generator=g)
    b0 = torch.randn(32, device=DEV, generator=g)
    h_all = torch.empty((U, 32), device=DEV)
    h_ids = torch.empty((1000, 32), device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    ops.call("gngf_hpd_first_layer_fwd", lat, w0.data_ptr(), b0.data_ptr(), 32, 1, h_all.data_ptr(), st)
    ops.call("gngf_hpd_first_layer_fwd_nodes", lat, ids.data_ptr(), 1000, w0.data_ptr(), b0.data_ptr(), 32, 1,
             h_ids.data_ptr(), st)
    assert torch.equal(h_ids, h_all[ids.long()])
    dz = torch.randn((1000, 32), device=DEV, generator=g)
    dz_all = torch.zeros((U, 32), device=DEV)
    dz_all[ids.long()] = dz
    dw_a, db_a, dw_b, db_b = (torch.zeros((32, 2), device=DEV), torch.zeros(32, device=DEV),
                              torch.zeros((32, 2), device=DEV), torch.zeros(32, device=DEV))
    ops.call("gngf_hpd_first_layer_bwd", lat, dz_all.data_ptr(), 32, dw_a.data_ptr(), db_a.data_ptr(), st)
    ops.call("gngf_hpd_first_layer_bwd_nodes", lat, ids.data_ptr(), 1000, dz.data_ptr(), 32, dw_b.data_ptr(),
             db_b.data_ptr(), st)
    assert rel_err(dw_b.cpu().numpy(), dw_a.cpu().numpy()) < 1e-5
    assert rel_err(db_b.cpu().numpy(), db_a.cpu().numpy()) < 1e-5


def test_corners_empty_batch():
    scaled, grid = ops.corners_fwd(torch.empty((0, 2), device=DEV), build_lattice([8, 16]))
    assert scaled.shape == (0, 2, 2, 1) and grid.shape == (0, 2, 2, 4)


@pytest.mark.parametrize("T", [256, 2 ** 14, 2 ** 19, 2 ** 22, 1000])
def test_fast_hash_bit_exact(T):
    n_ls = level_resolutions(16, 8192, 16)
    x = _edge_coords()
    idx = ops.fast_hash_fwd(torch.from_numpy(x).to(DEV), build_lattice(n_ls), T).cpu().numpy()
    _, grid = O.scale_to_grid(x, n_ls)
    assert np.array_equal(idx, O.fast_hash(grid.astype(np.int32), T))


def test_fast_hash_matches_reference_golden():
    g = load("hash_mode")
    idx = ops.fast_hash_fwd(torch.from_numpy(g["x"]).to(DEV), build_lattice(g["n_ls"]), g["cfg"]["T"])
    assert idx.dtype == torch.int64
    assert np.array_equal(idx.cpu().numpy(), g["idx"])


def _no_tie_rows(p, k, tol=0.0):
    srt = -np.sort(-p, axis=-1)
    return (srt[:, :k] - srt[:, 1:k + 1]).min(axis=-1) > tol


@pytest.mark.parametrize("R,T,K", [(1000, 256, 4), (77, 256, 1), (64, 256, 20), (33, 256, 128), (50, 1000, 4),
                                   (20, 4096, 32), (6, 2 ** 14, 4), (3, 2 ** 19, 4), (1, 7, 7)])
def test_softmax_topk_matches_oracle(R, T, K):
    rng = np.random.default_rng(R * 31 + T)
    logits = (rng.standard_normal((R, T)) * 3).astype(np.float32)
    probs, topv, topi = ops.softmax_topk_fwd(torch.from_numpy(logits).to(DEV), K)
    p_ref = O.softmax_lastdim(logits)
    assert rel_err(probs.cpu().numpy(), p_ref) < 1e-5
    # indices: bit-exact against the oracle's top-k *on the same probabilities* (ties -> lower index)
    p_gpu = probs.cpu().numpy()
    v_ref, i_ref = O.topk_sorted(p_gpu, K)
    assert np.array_equal(topi.cpu().numpy(), i_ref)
    assert np.array_equal(topv.cpu().numpy(), v_ref)
    # and against the oracle's own probabilities wherever its k-boundary is not a rounding-level tie
    ok = _no_tie_rows(p_ref, min(K, T - 1), tol=4e-7 * p_ref.max())
    assert np.array_equal(topi.cpu().numpy()[ok], O.topk_sorted(p_ref, K)[1][ok])


def test_softmax_topk_ties_nan_and_inplace():
    logits = np.zeros((4, 256), dtype=np.float32)                  # all ties -> indices 0..K-1
    logits[1, 5] = logits[1, 200] = 3.0                            # two-way tie at the top
    logits[2, :] = np.nan                                          # softmax -> NaN -> nan_to_num -> 0
    logits[3, 17] = np.inf                                         # exp(inf - inf) = NaN row as well
    t = torch.from_numpy(logits).to(DEV)
    probs, topv, topi = ops.softmax_topk_fwd(t, 4, inplace=True)
    assert probs.data_ptr() == t.data_ptr()
    topi = topi.cpu().numpy()
    assert list(topi[0]) == [0, 1, 2, 3]
    assert list(topi[1]) == [5, 200, 0, 1]
    assert list(topi[2]) == [0, 1, 2, 3] and float(probs[2].abs().sum()) == 0.0
    assert torch.isfinite(probs).all()


def test_topk_forward_backward_generic_values():
    rng = np.random.default_rng(3)
    v = rng.standard_normal((5, 7, 300)).astype(np.float32)
    v[0, 0, :10] = -np.inf
    v[1, 2, 3] = v[1, 2, 9] = 10.0
    v[2] = np.round(v[2])                                           # many duplicates
    t = torch.from_numpy(v).to(DEV).requires_grad_()
    from collision_handling_in_instantngp_b200.models import DifferentiableTopk
    vals, idx = DifferentiableTopk.apply(t, 6, -1)
    v_ref, i_ref = O.topk_sorted(v, 6)
    assert idx.dtype == torch.int64
    assert np.array_equal(idx.cpu().numpy(), i_ref) and np.array_equal(vals.detach().cpu().numpy(), v_ref)
    gv = torch.from_numpy(rng.standard_normal(v_ref.shape).astype(np.float32)).to(DEV)
    vals.backward(gv)
    g_ref = np.zeros_like(v)
    np.put_along_axis(g_ref, i_ref, gv.cpu().numpy(), axis=-1)      # models.py:27-35
    assert np.array_equal(t.grad.cpu().numpy(), g_ref)
    # another dim
    vals2, idx2 = DifferentiableTopk.apply(torch.from_numpy(v).to(DEV), 2, 1)
    v2, i2 = O.topk_sorted(np.moveaxis(v, 1, -1), 2)
    assert np.array_equal(np.moveaxis(idx2.cpu().numpy(), 1, -1), i2)


@pytest.mark.parametrize("M,N,K,act", [(1000, 64, 8, 1), (777, 3, 64, 3), (130, 256, 128, 0), (65, 33, 17, 2), (1, 1, 1, 1)])
def test_linear_forward_backward(M, N, K, act):
    rng = np.random.default_rng(M + N)
    x = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    xt, wt, bt = (torch.from_numpy(a).to(DEV) for a in (x, w, b))
    y = ops.linear_fwd(xt, wt, bt, act)
    z = x.astype(np.float64) @ w.astype(np.float64).T + b
    ref = {0: z, 1: np.maximum(z, 0), 2: np.where(z > 0, z, 0.01 * z), 3: 1 / (1 + np.exp(-z))}[act]
    assert rel_err(y.cpu().numpy(), ref) < 1e-5
    dz = rng.standard_normal((M, N)).astype(np.float32)
    dw, db = torch.zeros_like(wt), torch.zeros_like(bt)
    xin = np.maximum(x, 0)                                           # pretend x is a ReLU output
    dx = ops.linear_bwd(torch.from_numpy(dz).to(DEV), torch.from_numpy(xin).to(DEV), wt, 1, True, dw, db)
    assert rel_err(dw.cpu().numpy(), dz.astype(np.float64).T @ xin) < 1e-5
    assert rel_err(db.cpu().numpy(), dz.astype(np.float64).sum(0)) < 1e-5
    assert rel_err(dx.cpu().numpy(), (dz.astype(np.float64) @ w) * (xin > 0)) < 1e-5


@pytest.mark.parametrize("M,N,K", [(20001, 64, 32), (33000, 128, 64), (16384, 32, 64), (40007, 128, 128), (17000, 64, 64)])
def test_skinny_linear_layers(M, N, K):
    """The persistent skinny-layer kernels (k2_linear.cu: weights resident in shared memory, 128-row tiles; one-pass dW / db)
    that serve the hidden HPD layers on large lattices: forward, dX with the ReLU mask, dW, db against float64."""
    g = torch.Generator(device=DEV).manual_seed(M + N + K)
    x = torch.randn((M, K), device=DEV, generator=g)
    w = torch.randn((N, K), device=DEV, generator=g) / K ** 0.5
    b = torch.randn(N, device=DEV, generator=g)
    y = ops.linear_fwd(x, w, b, 1)
    ref = torch.relu(x.double() @ w.double().T + b.double())
    assert float((y.double() - ref).abs().max()) < 1e-5 * float(ref.abs().max())
    dz = torch.randn((M, N), device=DEV, generator=g)
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    xin = torch.relu(x)
    dx = ops.linear_bwd(dz, xin, w, 1, True, dw, db)
    assert float((dw.double() - dz.double().T @ xin.double()).abs().max()) < 1e-5 * float(dw.abs().max())
    assert float((db.double() - dz.double().sum(0)).abs().max()) < 1e-5 * float(db.abs().max())
    assert float((dx.double() - (dz.double() @ w.double()) * (xin > 0)).abs().max()) < 1e-5 * float(dx.abs().max())
    # dW only / dX only (the call shapes of the small-lattice path)
    dw2 = torch.zeros_like(w)
    assert ops.linear_bwd(dz, xin, w, 0, False, dw2, None) is None
    assert float((dw2 - dw).abs().max()) < 1e-5 * float(dw.abs().max())


def test_linear_layers_with_more_than_65535_row_tiles():
    """67 M lattice nodes (the 8192^2 configuration) are > 65 535 row tiles of 64: the row tiles sit on grid.x."""
    M, N, K = 64 * 70000 + 7, 8, 4
    g = torch.Generator(device=DEV).manual_seed(1)
    x = torch.randn((M, K), device=DEV, generator=g)
    w = torch.randn((N, K), device=DEV, generator=g)
    b = torch.randn(N, device=DEV, generator=g)
    y = ops.linear_fwd(x, w, b, 1)
    ref = torch.relu(x.double() @ w.double().T + b.double())
    assert float((y.double() - ref).abs().max()) < 1e-5 * float(ref.abs().max())
    dz = torch.randn((M, N), device=DEV, generator=g)
    dw, db = torch.zeros_like(w), torch.zeros_like(b)
    xin = torch.relu(x)
    dx = ops.linear_bwd(dz, xin, w, 1, True, dw, db)
    assert float((dw.double() - dz.double().T @ xin.double()).abs().max()) < 1e-4 * float(dw.abs().max())
    assert float((db.double() - dz.double().sum(0)).abs().max()) < 1e-4 * float(db.abs().max())
    assert float((dx.double() - (dz.double() @ w.double()) * (xin > 0)).abs().max()) < 1e-5 * float(dx.abs().max())


def test_missing_gpu_tensor_is_rejected():
    from collision_handling_in_instantngp_b200 import GngfError
    with pytest.raises(GngfError):
        ops.corners_fwd(torch.zeros(4, 2), build_lattice([8]))


@pytest.mark.parametrize("P,IN,OUT,leaky", [(1000, 8, 3, 0), (57404, 8, 3, 0), (333, 32, 3, 1), (129, 6, 1, 0),
                                             (4097, 64, 3, 0), (1, 8, 3, 0)])
def test_fused_decoder_mlp(P, IN, OUT, leaky):
    """K6 against a float64 evaluation of models.py:382-392 and its gradients."""
    from collision_handling_in_instantngp_b200 import _lib
    rng = np.random.default_rng(P + IN)
    enc = (rng.standard_normal((P, IN)) * 0.5).astype(np.float32)
    ws = [(rng.standard_normal((o, i)) / np.sqrt(i)).astype(np.float32) for i, o in [(IN, 64), (64, 64), (64, OUT)]]
    bs = [(rng.standard_normal(o) * 0.1).astype(np.float32) for o in (64, 64, OUT)]
    drgb = rng.standard_normal((P, OUT)).astype(np.float32)
    t = lambda a: torch.from_numpy(a).to(DEV)
    enc_t, ws_t, bs_t, drgb_t = t(enc), [t(w) for w in ws], [t(b) for b in bs], t(drgb)
    rgb = torch.empty((P, OUT), device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    _lib.call("gngf_mlp3_fwd", enc_t.data_ptr(), P, IN, OUT, leaky, ws_t[0].data_ptr(), bs_t[0].data_ptr(),
              ws_t[1].data_ptr(), bs_t[1].data_ptr(), ws_t[2].data_ptr(), bs_t[2].data_ptr(), rgb.data_ptr(), st)
    acts = O.decoder_forward(enc.astype(np.float64), [w.astype(np.float64) for w in ws],
                             [b.astype(np.float64) for b in bs], bool(leaky))
    assert rel_err(rgb.cpu().numpy(), acts[-1]) < 1e-5
    denc = torch.empty((P, IN), device=DEV)
    gw = [torch.zeros_like(w) for w in ws_t]
    gb = [torch.zeros_like(b) for b in bs_t]
    work = torch.empty(_lib.load().gngf_mlp3_bwd_workspace_floats(IN, OUT), device=DEV)
    _lib.call("gngf_mlp3_bwd", enc_t.data_ptr(), drgb_t.data_ptr(), P, IN, OUT, leaky, ws_t[0].data_ptr(),
              bs_t[0].data_ptr(), ws_t[1].data_ptr(), bs_t[1].data_ptr(), ws_t[2].data_ptr(), bs_t[2].data_ptr(),
              denc.data_ptr(), gw[0].data_ptr(), gb[0].data_ptr(), gw[1].data_ptr(), gb[1].data_ptr(), gw[2].data_ptr(),
              gb[2].data_ptr(), work.data_ptr(), st)
    dz = drgb.astype(np.float64) * acts[-1] * (1 - acts[-1])
    dws, dbs, dx = O._mlp_backward(acts[:-1], [w.astype(np.float64) for w in ws], dz, leaky=bool(leaky))
    assert rel_err(denc.cpu().numpy(), dx) < 1e-5
    for i in range(3):
        assert rel_err(gw[i].cpu().numpy(), dws[i]) < 2e-5, i
        assert rel_err(gb[i].cpu().numpy(), dbs[i]) < 2e-5, i


@pytest.mark.parametrize("P,IN,OUT,leaky", [(1000, 8, 3, 0), (57404, 8, 3, 0), (333, 32, 3, 1), (129, 6, 1, 0),
                                             (4097, 64, 3, 0), (1, 8, 3, 0), (40000, 32, 3, 0), (700, 20, 4, 1)])
def test_tensor_core_decoder_mlp(P, IN, OUT, leaky):
    """k6_mlp_tc.cu (tcgen05 chains, operands as bf16 planes kept in shared memory) against float64.
    Forward: three planes, bar 1e-5; backward: two planes, bar 1e-4 (gradients accumulate into the given buffers).
    The backward gates with the ReLU pattern the forward recorded; the float64 reference differentiates that same
    pattern (it may differ from float64's own on units whose pre-activation is within the forward's rounding error
    of zero -- asserted to be a < 1e-5 fraction)."""
    from collision_handling_in_instantngp_b200 import _lib
    rng = np.random.default_rng(P + IN)
    enc = (rng.standard_normal((P, IN)) * 0.5).astype(np.float32)
    ws = [(rng.standard_normal((o, i)) / np.sqrt(i)).astype(np.float32) for i, o in [(IN, 64), (64, 64), (64, OUT)]]
    bs = [(rng.standard_normal(o) * 0.1).astype(np.float32) for o in (64, 64, OUT)]
    drgb = rng.standard_normal((P, OUT)).astype(np.float32)
    t = lambda a: torch.from_numpy(a).to(DEV)
    enc_t, ws_t, bs_t, drgb_t = t(enc), [t(w) for w in ws], [t(b) for b in bs], t(drgb)
    rgb = torch.empty((P, OUT), device=DEV)
    masks_t = torch.empty((P, 4), dtype=torch.int32, device=DEV)
    st = torch.cuda.current_stream().cuda_stream
    _lib.call("gngf_mlp3_tc_fwd", enc_t.data_ptr(), P, IN, OUT, leaky, ws_t[0].data_ptr(), bs_t[0].data_ptr(),
              ws_t[1].data_ptr(), bs_t[1].data_ptr(), ws_t[2].data_ptr(), bs_t[2].data_ptr(), rgb.data_ptr(),
              masks_t.data_ptr(), st)
    w64, b64 = [w.astype(np.float64) for w in ws], [b.astype(np.float64) for b in bs]
    acts = O.decoder_forward(enc.astype(np.float64), w64, b64, bool(leaky))
    assert rel_err(rgb.cpu().numpy(), acts[-1]) < 1e-5
    m = masks_t.cpu().numpy().view(np.uint32)
    bits = ((m[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).astype(bool)
    masks = [None, bits[:, 0:2].reshape(P, 64), bits[:, 2:4].reshape(P, 64)]
    for i in (1, 2):
        assert (masks[i] != (acts[i] > 0)).mean() < 1e-5
    denc = torch.empty((P, IN), device=DEV)
    gw0 = [rng.standard_normal(w.shape).astype(np.float32) for w in ws]      # accumulated into
    gb0 = [rng.standard_normal(b.shape).astype(np.float32) for b in bs]
    gw, gb = [t(a) for a in gw0], [t(a) for a in gb0]
    _lib.call("gngf_mlp3_tc_bwd", enc_t.data_ptr(), rgb.data_ptr(), drgb_t.data_ptr(), P, IN, OUT, leaky,
              ws_t[0].data_ptr(), bs_t[0].data_ptr(), ws_t[1].data_ptr(), bs_t[1].data_ptr(), ws_t[2].data_ptr(),
              masks_t.data_ptr(), denc.data_ptr(), gw[0].data_ptr(), gb[0].data_ptr(), gw[1].data_ptr(),
              gb[1].data_ptr(), gw[2].data_ptr(), gb[2].data_ptr(), st)
    dz = drgb.astype(np.float64) * acts[-1] * (1 - acts[-1])
    dws, dbs, dx = O._mlp_backward(acts[:-1], w64, dz, leaky=bool(leaky), masks=masks)
    assert rel_err(denc.cpu().numpy(), dx) < 1e-4
    for i in range(3):
        # the random starting contents limit what float32 accumulation can resolve: compare the increments
        scale = max(np.abs(dws[i]).max(), 1.0)
        assert np.abs(gw[i].cpu().numpy().astype(np.float64) - gw0[i] - dws[i]).max() < 1e-4 * scale, i
        scale = max(np.abs(dbs[i]).max(), 1.0)
        assert np.abs(gb[i].cpu().numpy().astype(np.float64) - gb0[i] - dbs[i]).max() < 1e-4 * scale, i


def test_generic_decoder_shape_uses_linear_layers():
    """A decoder that is not IN-64-64-OUT goes through gngf_linear_fwd/bwd and still matches the oracle."""
    from golden_util import loss_cfg, oracle_cfg, params_of
    from parity_util import build_net, run_step
    g = dict(load("cfg2_small"))
    rng = np.random.default_rng(5)
    g["cfg"] = dict(g["cfg"], mlp=[32, 48])
    shapes = {"mlp.0.0": (32, 8), "mlp.1.0": (48, 32), "mlp.2.0": (3, 48)}
    for k, (o, i) in shapes.items():
        g[f"param.{k}.weight"] = (rng.standard_normal((o, i)) / np.sqrt(i)).astype(np.float32)
        g[f"param.{k}.bias"] = (rng.standard_normal(o) * 0.1).astype(np.float32)
    net = build_net(g)
    out = run_step(net, g)
    assert not out["state"].mlp_fused
    p = params_of(g)
    fwd = O.gngf_forward(p, g["x"], oracle_cfg(g))
    grads = O.gngf_backward(p, g["x"], g["y"], oracle_cfg(g), fwd, loss_cfg(g))
    assert rel_err(out["rgb"], fwd["rgb"]) < 1e-5
    for i in range(3):
        assert rel_err(out["grads"][f"mlp.{i}.0.weight"], grads["mlp_w"][i]) < 1e-4
    assert rel_err(out["grads"]["encoding._hash_tables.2.weight"], grads["tables"][2]) < 1e-4


@pytest.mark.parametrize("M,N,K,act", [(128, 128, 128, 0), (300, 256, 128, 0), (1000, 1000, 64, 1), (782, 256, 128, 0),
                                        (4096, 8192, 128, 0), (130, 136, 192, 0), (5, 8, 8, 0)])
def test_tensor_core_gemm_matches_float64(M, N, K, act):
    """tcgen05 split-bf16 GEMM: fp32-level accuracy (the bar is 1e-5 relative on the logits)."""
    rng = np.random.default_rng(M + N + K)
    x = (rng.standard_normal((M, K)) * 2).astype(np.float32)
    x[x < -1] = 0                                                   # ReLU-like sparsity, as h3 has
    w = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    y = ops.tc_linear_fwd(torch.from_numpy(x).to(DEV), torch.from_numpy(w).to(DEV), torch.from_numpy(b).to(DEV), act)
    ref = x.astype(np.float64) @ w.astype(np.float64).T + b
    if act == 1:
        ref = np.maximum(ref, 0)
    err = rel_err(y.cpu().numpy(), ref)
    # measured on B200: 1.2e-6 .. 1.8e-6 (the TMEM accumulator does not round like an IEEE FMA chain, whose
    # error on the same inputs is 3e-7 .. 5e-7); plain bf16 operands would give ~4e-3
    assert err < 5e-6, err


def test_split_bf16x3_reconstructs_fp32():
    rng = np.random.default_rng(0)
    x = (rng.standard_normal(100000) * np.exp(rng.uniform(-20, 20, 100000))).astype(np.float32)
    planes = ops.split_bf16x3(torch.from_numpy(x).to(DEV))
    rec = planes.double().sum(0).cpu().numpy()
    assert np.abs(rec - x).max() <= np.abs(x).max() * 2.0 ** -22
    assert np.all(np.abs(rec - x) <= np.abs(x) * 2.0 ** -21)


@pytest.mark.parametrize("U,T,Kd,K", [(300, 1000, 128, 4), (782, 256, 128, 4), (130, 4096, 64, 1), (1000, 2 ** 14, 128, 8),
                                       (64, 2 ** 17, 128, 4), (5000, 8192, 128, 4)])
def test_streaming_hpd_matches_unfused(U, T, Kd, K):
    """Fused tcgen05 output layer + online softmax + running top-k == GEMM -> softmax/top-k kernel."""
    rng = np.random.default_rng(U + T)
    h = np.maximum(rng.standard_normal((U, Kd)), 0).astype(np.float32)
    w = (rng.standard_normal((T, Kd)) * (3.0 / np.sqrt(Kd))).astype(np.float32)
    b = rng.standard_normal(T).astype(np.float32)
    ht, wt, bt = (torch.from_numpy(a).to(DEV) for a in (h, w, b))
    utopv, utopi, rmax, rsum = ops.hpd_stream_fwd(ht, wt, bt, K)
    logits = h.astype(np.float64) @ w.astype(np.float64).T + b
    p_ref = O.softmax_lastdim(logits)
    v_ref, i_ref = O.topk_sorted(p_ref, K)
    srt = -np.sort(-p_ref, axis=-1)
    gap = (srt[:, :K] - srt[:, 1:K + 1]).min(-1) / srt[:, 0]
    ok = gap > 2e-5                                                  # rows whose selection is not a rounding-level tie
    # measured rate of such ties on these inputs: below 1e-3 of the rows per selected slot (0 / 300, 2 / 782, 6 / 1000 at
    # K = 8: K + 1 order statistics of T logits per row, each pair a chance of a 2e-5 gap)
    print(f"rows exempt as rounding-level ties: {int((~ok).sum())} of {U}")
    assert (~ok).sum() <= max(2, 0.002 * U * K)
    assert np.array_equal(utopi.cpu().numpy()[ok], i_ref[ok])
    assert rel_err(utopv.cpu().numpy()[ok], v_ref[ok]) < 1e-5
    assert rel_err(rmax.cpu().numpy(), logits.max(-1)) < 1e-5
    assert rel_err(rsum.cpu().numpy(), np.exp(logits - logits.max(-1, keepdims=True)).sum(-1)) < 1e-5
    # every selected index is a true top-k member up to the tie tolerance, on all rows
    sel = np.take_along_axis(p_ref, utopi.cpu().numpy().astype(np.int64), -1)
    assert (sel >= srt[:, K - 1:K] * (1 - 2e-5)).all()


@pytest.mark.parametrize("regime", ["unit_scale_inputs", "raw_init"])
def test_streaming_hpd_at_the_size_of_configs2(regime):
    """The streaming path at the size BASELINE.json configs[2] names -- 173 400 lattice nodes (macaw lattice, 16 levels)
    x T = 2^19 slots: 4 096 column tiles, Kahan row sums over 5e5 terms, the column-split merge -- against a float64
    evaluation of EVERY row (torch fp64 on the GPU, row chunks): selections, probabilities, softmax statistics, and
    the three gradients of the fused backward (dh, dW3, db3) at full size, on the real activations of a random-init HPD.

    The HPD's input is the integer lattice coordinate (models.py:416-418; up to 509 here), so with nn.Linear's initial
    weights the logits are O(1e2..1e3): exp() turns their fp32 rounding error (eps * |z| * sqrt(128)) into a relative
    error of that size in every probability, and no fp32 evaluation -- the reference's included -- reproduces fp64 to
    1e-5 / 1e-4 there.  Hence two regimes:
      unit_scale_inputs  first layer scaled by 1/512, logits O(1): the stated bars -- indices exact on every row an fp32
                         evaluation can decide (the exempt rows are counted), probabilities / row sums 1e-5, gradients 1e-4;
      raw_init           the same quantities next to the error of a plain fp32 torch evaluation of the same rows
                         (allow_tf32 off: what the reference computes).  Bounds: forward 1e-5 or 8 x that fp32 noise;
                         dW3 / db3 1e-4; dh 2e-4 -- measured 1.5e-4: the logits are recomputed from two fp16 planes
                         (22 bits per operand), and with |h| ~ 300 the dot products' absolute error (~2e-4) is the
                         relative error of every unselected probability that enters dh; the selected slots, which
                         dominate, are exact fp32 products (see k2_hpd_tc_bwd.cu)."""
    from collision_handling_in_instantngp_b200.models import GeneralNeuralGaugeFields
    torch.manual_seed(65535)
    T, K, Kd = 2 ** 19, 4, 128
    net = GeneralNeuralGaugeFields(input_dim=2, hash_table_size=T, num_levels=16, n_min=16, n_max=508,
                                   MLP_hidden_layers_widths=[64, 64], HPD_hidden_layers_widths=[32, 64, 128],
                                   HPD_out_features=T, topk_k=K, should_keep_topk_only=True)
    lat = build_lattice(level_resolutions(16, 508, 16), (0.0, 0.0), (1.0, 338 / 507))
    U = lat.num_nodes
    assert U == 173400
    ws, bs = net.HPD.weights()
    with torch.no_grad():
        if regime == "unit_scale_inputs":
            ws[0].mul_(1.0 / 512)
        # the real activations of the last hidden layer on the real lattice (random-init HPD, nn.Linear bounds)
        h = torch.empty((U, 32), device=DEV)
        ops.call("gngf_hpd_first_layer_fwd_nodes", lat, None, U, ws[0].data_ptr(), bs[0].data_ptr(), 32, ops.ACT_RELU,
                 h.data_ptr(), ops._stream())
        for i in (1, 2):
            h = ops.linear_fwd(h, ws[i], bs[i], ops.ACT_RELU)
        w, b = ws[3].detach().contiguous(), bs[3].detach().contiguous()
    g = torch.Generator(device=DEV).manual_seed(7)
    dtv = torch.randn((U, K), generator=g, device=DEV)
    hp, wp = ops.split_f16x2(h), ops.split_f16x2(w)
    utopv, utopi, rmax, rsum = ops.hpd_stream_fwd(h, w, b, K, h_planes=hp, w_planes=wp)
    dw = torch.zeros((T, Kd), device=DEV)
    db = torch.zeros(T, device=DEV)
    dh = ops.hpd_stream_bwd(_flat_lattice(U), h, w, b, hp, wp, utopv, utopi, dtv, None, None, rmax, rsum, dw, db)
    torch.cuda.synchronize()
    del hp, wp

    w64, b64 = w.double(), b.double()
    dw64 = torch.zeros((T, Kd), dtype=torch.float64, device=DEV)
    db64 = torch.zeros(T, dtype=torch.float64, device=DEV)
    CH = 1024
    EPS = float(np.finfo(np.float32).eps)
    m_all = torch.empty(U, dtype=torch.float64, device=DEV)
    s_all = torch.empty(U, dtype=torch.float64, device=DEV)
    p_all = torch.empty((U, K), dtype=torch.float64, device=DEV)
    dh_all = torch.empty((U, Kd), dtype=torch.float64, device=DEV)
    wrong, undecidable, zmax = 0, 0, 0.0
    worst = dict(topv=0.0, rsum=0.0, rmax=0.0, dh=0.0)
    noise32 = dict(topv=0.0, rsum=0.0, dh=0.0)
    dh_scale = float(dh.abs().max())
    for r0 in range(0, U, CH):
        r1 = min(U, r0 + CH)
        h64 = h[r0:r1].double()
        z = torch.addmm(b64, h64, w64.t())                                   # (rows, T) fp64
        zt, it = z.topk(K + 1, dim=-1)
        m = zt[:, 0:1]
        zmax = max(zmax, float(zt.abs().max()))
        if r0 == 0:
            # what a plain fp32 evaluation (the reference's arithmetic) makes of the same rows
            z32 = torch.addmm(b, h[r0:r1], w.t())
            m32 = z32.max(-1, keepdim=True).values
            e32 = (z32 - m32).exp()
            s32 = e32.sum(-1, keepdim=True)
            ti0 = utopi[r0:r1].long()
            p32 = e32.gather(1, ti0) / s32
            pg32 = p32 * dtv[r0:r1]
            dl32 = e32 * (-pg32.sum(-1, keepdim=True) / s32)
            dl32.scatter_add_(1, ti0, pg32)
            dh32 = (dl32 @ w) * (h[r0:r1] > 0)
            del z32, e32, dl32
        z.sub_(m).exp_()                                                     # e = exp(z - max), in place
        ssum = z.sum(-1, keepdim=True)
        ti = utopi[r0:r1].long()
        same = (ti == it[:, :K]).all(-1)
        gap = (zt[:, :-1] - zt[:, 1:]).min(-1).values
        # a selection is decidable in fp32 when the K+1 best logits are further apart than the rounding error of a
        # 128-term fp32 dot product of their size (64 eps |z|: ~6 x the typical error of one evaluation)
        decidable = gap > 64 * EPS * zt[:, 0].abs().clamp_min(1.0)
        wrong += int((~same & decidable).sum())
        undecidable += int((~decidable).sum())
        p_sel = z.gather(1, ti) / ssum                                       # probabilities of OUR selection
        m_all[r0:r1], s_all[r0:r1], p_all[r0:r1] = m[:, 0], ssum[:, 0], p_sel
        pv_ref = zt[:, :K].sub(m).exp() / ssum
        tv_err = float((utopv[r0:r1].double() - pv_ref)[same].abs().max() / pv_ref.max())
        worst["topv"] = max(worst["topv"], tv_err)
        worst["rsum"] = max(worst["rsum"], float(((rsum[r0:r1].double() - ssum[:, 0]).abs() / ssum[:, 0]).max()))
        worst["rmax"] = max(worst["rmax"], float(((rmax[r0:r1].double() - m[:, 0]).abs() / m[:, 0].abs().clamp_min(1.0)).max()))
        # backward in fp64, differentiating the kernel's own selection: dl = -<g, p_sel> p + scatter(p_sel g)
        pg = p_sel * dtv[r0:r1].double()
        z.mul_(-pg.sum(-1, keepdim=True) / ssum)                             # dl, in place
        z.scatter_add_(1, ti, pg)
        dw64.addmm_(z.t(), h64)
        db64.add_(z.sum(0))
        dh64 = (z @ w64) * (h64 > 0)
        dh_all[r0:r1] = dh64
        worst["dh"] = max(worst["dh"], float((dh[r0:r1].double() - dh64).abs().max()) / dh_scale)
        if r0 == 0:
            noise32["topv"] = float((p32.double() - p_sel).abs().max() / p_sel.max())
            noise32["rsum"] = float(((s32[:, 0].double() - ssum[:, 0]).abs() / ssum[:, 0]).max())
            noise32["dh"] = float((dh32.double() - dh64).abs().max()) / dh_scale
        del z, dh64
    worst["dw"] = float((dw.double() - dw64).abs().max() / dw64.abs().max())
    worst["db"] = float((db.double() - db64).abs().max() / db64.abs().max())
    # the backward kernels alone: the same call fed with the float64 softmax statistics (rounded to fp32) instead of the
    # forward's -- separates the backward's own error from what it inherits through row_max / row_sum / utopv
    hp, wp = ops.split_f16x2(h), ops.split_f16x2(w)
    dw2 = torch.zeros((T, Kd), device=DEV)
    db2 = torch.zeros(T, device=DEV)
    dh2 = ops.hpd_stream_bwd(_flat_lattice(U), h, w, b, hp, wp, p_all.float(), utopi, dtv, None, None, m_all.float(),
                             s_all.float(), dw2, db2)
    torch.cuda.synchronize()
    alone = {"dh": float((dh2.double() - dh_all).abs().max()) / dh_scale,
             "dw": float((dw2.double() - dw64).abs().max() / dw64.abs().max()),
             "db": float((db2.double() - db64).abs().max() / db64.abs().max())}
    del hp, wp, dh_all
    print(f"\n[{regime}] configs[2] size (U = {U}, T = 2^19, |logit| up to {zmax:.3g}): rows with a wrong selection {wrong}, rows "
          f"no fp32 evaluation can decide {undecidable}; worst errors vs fp64 {({k: float(f'{v:.2e}') for k, v in worst.items()})}; "
          f"a plain fp32 evaluation of the first {CH} rows: {({k: float(f'{v:.2e}') for k, v in noise32.items()})}; the backward "
          f"alone (fed the fp64 statistics): {({k: float(f'{v:.2e}') for k, v in alone.items()})}")
    assert wrong == 0
    # (with logits of |z| <= 0.26 over 5e5 slots the K + 1 best are often closer than fp32 can resolve: 1.2 % of the rows)
    assert undecidable <= (0.02 if regime == "unit_scale_inputs" else 0.002) * U
    if regime == "unit_scale_inputs":
        fwd_bar = {k: 1e-5 for k in ("topv", "rsum", "rmax")}
        grad_bar = {"dh": 1e-4, "dw": 1e-4, "db": 1e-4}
    else:
        fwd_bar = {"topv": max(1e-5, 8 * noise32["topv"]), "rsum": max(1e-5, 8 * noise32["rsum"]), "rmax": 1e-6}
        grad_bar = {"dh": 2e-4, "dw": 1e-4, "db": 1e-4}
    for k, bar in fwd_bar.items():
        assert worst[k] < bar, (k, worst[k], bar)
    for k, bar in grad_bar.items():
        assert worst[k] < bar and alone[k] < bar, (k, worst[k], alone[k], bar)


def test_zero_skipping_changes_nothing(monkeypatch):
    """The streaming kernels skip what a one-hot softmax makes exactly zero (dead 32-column chunks, all-zero E tiles:
    k2_hpd_tc_bwd.cu, k2_hpd_tc.cu).  With GNGF_DEBUG_NO_SKIP=1 they take the full path on every chunk and tile.  On a
    random-init HPD fed integer coordinates up to 8 191 (BASELINE.json configs[3]: logits O(1e4), > 95 % of the tiles
    dead) the two must agree: selections and row maxima bit for bit, row sums / probabilities to 1e-6 (a skipped chunk
    also skips a no-op step of the compensated sum), and -- fed the same forward outputs -- dh BIT FOR BIT (one CTA owns
    a row tile and walks its column tiles in order) and dW3 / db3 to the reordering noise of their atomics (1e-6 of the
    largest entry); the fast run must indeed have skipped most second products (gngf_hpd_stream_bwd_stats)."""
    from collision_handling_in_instantngp_b200.models import GeneralNeuralGaugeFields
    torch.manual_seed(65535)
    T, K, Kd = 2 ** 14, 4, 128
    net = GeneralNeuralGaugeFields(input_dim=2, hash_table_size=T, num_levels=16, n_min=16, n_max=8192,
                                   MLP_hidden_layers_widths=[64, 64], HPD_hidden_layers_widths=[32, 64, 128],
                                   HPD_out_features=T, topk_k=K, should_keep_topk_only=True)
    ws, bs = net.HPD.weights()
    U = 40000
    g = torch.Generator(device=DEV).manual_seed(3)
    with torch.no_grad():
        # lattice coordinates: runs of neighbouring nodes at scattered places of the 8193 x 8193 lattice, plus the origin
        base = torch.randint(0, 8192, (U // 100, 2), generator=g, device=DEV)
        xy = (base[:, None, :] + torch.stack([torch.arange(100, device=DEV), torch.zeros(100, dtype=torch.long, device=DEV)], -1)[None])
        xy = xy.reshape(-1, 2).clamp_(0, 8192).float()
        xy[:200] = torch.stack([torch.arange(200, device=DEV).float() % 20, torch.arange(200, device=DEV).float() // 20], -1)
        h = torch.relu(xy @ ws[0].t() + bs[0])
        for i in (1, 2):
            h = ops.linear_fwd(h, ws[i], bs[i], ops.ACT_RELU)
        w, b = ws[3].detach().contiguous(), bs[3].detach().contiguous()
    dtv = torch.randn((U, K), generator=g, device=DEV)

    hp, wp = ops.split_f16x2(h), ops.split_f16x2(w)

    def forward():
        out = ops.hpd_stream_fwd(h, w, b, K, h_planes=hp, w_planes=wp)
        torch.cuda.synchronize()
        return out

    def backward(utopv, utopi, rmax, rsum):
        dw = torch.zeros((T, Kd), device=DEV)
        db = torch.zeros(T, device=DEV)
        dh = ops.hpd_stream_bwd(_flat_lattice(U), h, w, b, hp, wp, utopv, utopi, dtv, None, None, rmax, rsum, dw, db)
        torch.cuda.synchronize()
        return dh, dw, db

    ops.stream_bwd_stats(reset=True)
    fwd_fast = forward()
    bwd_fast = backward(*fwd_fast)
    st = ops.stream_bwd_stats(reset=True)
    monkeypatch.setenv("GNGF_DEBUG_NO_SKIP", "1")
    fwd_full = forward()
    bwd_full = backward(*fwd_fast)          # the SAME forward outputs: the backward passes are compared on their own
    st_full = ops.stream_bwd_stats(reset=True)
    monkeypatch.delenv("GNGF_DEBUG_NO_SKIP")
    print(f"\ntiles / with all three logit products / with the second product: dh pass {st[0]} / {st[1]} / {st[2]}, dW3 pass "
          f"{st[3]} / {st[4]} / {st[5]}; without skipping {st_full}")
    assert st_full[1] == st_full[0] and st_full[2] == st_full[0] and st_full[4] == st_full[3] and st_full[5] == st_full[3]
    assert st[2] <= st[1] < 0.5 * st[0] and st[5] <= st[4] < 0.5 * st[3] and st[2] < 0.2 * st[0] and st[5] < 0.2 * st[3]
    assert torch.equal(fwd_fast[1], fwd_full[1]) and torch.equal(fwd_fast[2], fwd_full[2])      # selections, row maxima
    assert float(((fwd_fast[3] - fwd_full[3]).abs() / fwd_full[3]).max()) < 1e-6                 # row sums
    assert float((fwd_fast[0] - fwd_full[0]).abs().max()) < 1e-6                                 # probabilities (<= 1)
    assert torch.equal(bwd_fast[0], bwd_full[0]), ("dh", float((bwd_fast[0] - bwd_full[0]).abs().max()))
    for name, a, c in zip(("dw", "db"), bwd_fast[1:], bwd_full[1:]):
        assert float((a - c).abs().max()) <= 1e-6 * float(c.abs().max()), name


def _flat_lattice(U):
    """A one-level lattice whose node box is U x 1 (the kernel-level tests need node ids only)."""
    from collision_handling_in_instantngp_b200._lib import Lattice
    lat = Lattice()
    lat.num_levels = 1
    lat.n[0] = max(U - 1, 1)
    lat.ox = lat.oy = lat.lox[0] = lat.loy[0] = 0
    lat.wx, lat.wy, lat.lwx[0], lat.lwy[0] = U, 1, U, 1
    lat.loff[0], lat.loff[1] = 0, U
    return lat


@pytest.mark.parametrize("U,T,Kd,K", [(300, 1000, 128, 4), (782, 256, 128, 4), (130, 4096, 64, 1), (1000, 2 ** 14, 128, 8),
                                       (64, 2 ** 17, 128, 4), (5000, 8192, 128, 4), (129, 72, 72, 2)])
def test_streaming_hpd_backward_matches_fp64(U, T, Kd, K):
    """Fused tcgen05 backward of the streaming output layer (logits recomputed in TMEM, dlogits never written; second
    product reads the streamed tile MN-major) against a float64 restatement.  Gradient bar: 1e-4 relative."""
    rng = np.random.default_rng(U * 7 + T)
    h = np.maximum(rng.standard_normal((U, Kd)), 0).astype(np.float32)
    w = (rng.standard_normal((T, Kd)) * (3.0 / np.sqrt(Kd))).astype(np.float32)
    b = rng.standard_normal(T).astype(np.float32)
    dtv = rng.standard_normal((U, K)).astype(np.float32)
    ht, wt, bt, dtvt = (torch.from_numpy(a).to(DEV) for a in (h, w, b, dtv))
    hp, wp = ops.split_f16x2(ht), ops.split_f16x2(wt)
    utopv, utopi, rmax, rsum = ops.hpd_stream_fwd(ht, wt, bt, K, h_planes=hp if K <= 4 else None, w_planes=wp if K <= 4 else None)
    dw0 = rng.standard_normal((T, Kd)).astype(np.float32)          # dw / db are accumulated into
    db0 = rng.standard_normal(T).astype(np.float32)
    dw, db = torch.from_numpy(dw0).to(DEV), torch.from_numpy(db0).to(DEV)
    dh = ops.hpd_stream_bwd(_flat_lattice(U), ht, wt, bt, hp, wp, utopv, utopi, dtvt, None, None, rmax, rsum, dw, db)
    torch.cuda.synchronize()

    logits = h.astype(np.float64) @ w.astype(np.float64).T + b
    p = O.softmax_lastdim(logits)
    ti = utopi.cpu().numpy().astype(np.int64)
    pk = np.take_along_axis(p, ti, -1)
    pg = pk * dtv
    dl = -pg.sum(-1, keepdims=True) * p
    np.add.at(dl, (np.arange(U)[:, None], ti), pg)
    dh_ref = (dl @ w.astype(np.float64)) * (h > 0)
    dw_ref = dl.T @ h.astype(np.float64)
    db_ref = dl.sum(0)
    assert rel_err(dh.cpu().numpy(), dh_ref) < 1e-4
    assert rel_err(dw.cpu().numpy() - dw0, dw_ref) < 1e-4
    assert rel_err(db.cpu().numpy() - db0, db_ref) < 1e-4


def test_fused_adam_matches_torch_adam():
    """k9_adam.cu (one launch for all tensors, device-side step counter) against torch.optim.Adam with the
    reference's settings (functions.py:96-127): per-group lr / weight decay, betas (0.9, 0.99), eps 1e-15."""
    from collision_handling_in_instantngp_b200.optim import FusedAdam
    rng = np.random.default_rng(3)
    shapes = [(256, 2), (256, 2), (64, 8), (64,), (64, 64), (3, 64), (3,), (1, 1), (1000, 37)]
    init = [rng.standard_normal(s).astype(np.float32) for s in shapes]

    def make(cls, **kw):
        ps = [torch.nn.Parameter(torch.from_numpy(a.copy()).to(DEV)) for a in init]
        opt = cls([{"params": ps[:2], "lr": 1e-4, "weight_decay": 0.0}, {"params": ps[2:5], "lr": 1e-3, "weight_decay": 1e-6},
                   {"params": ps[5:], "lr": 2e-3, "weight_decay": 1e-2}], betas=(0.9, 0.99), eps=1e-15, **kw)
        return ps, opt

    pa, oa = make(torch.optim.Adam)
    pb, ob = make(FusedAdam)
    for step in range(12):
        for i, (a, b) in enumerate(zip(pa, pb)):
            g = torch.from_numpy((rng.standard_normal(a.shape) * 10.0 ** rng.integers(-6, 1)).astype(np.float32)).to(DEV)
            if step == 5 and i == 7:
                g = None                                   # a parameter without a gradient is skipped
            a.grad = g
            b.grad = None if g is None else g.clone()
        oa.step()
        ob.step()
        for a, b in zip(pa, pb):
            assert rel_err(b.detach().cpu().numpy(), a.detach().cpu().numpy()) < 2e-6
    assert ob.step_count(pb[0]) == 12 and ob.step_count(pb[7]) == 11
    # state dicts are interchangeable with torch.optim.Adam's (the reference saves whole_opt.pt, functions.py:768):
    # continue the FusedAdam run in a stock Adam and vice versa, then take one more step in each
    pc, oc = make(torch.optim.Adam)
    pd, od = make(FusedAdam)
    import copy
    oc.load_state_dict(copy.deepcopy(ob.state_dict()))      # (load_state_dict aliases same-device tensors: copy, as a
    od.load_state_dict(copy.deepcopy(oa.state_dict()))      #  torch.save / torch.load round trip would)
    for src, dst in ((pb, pc), (pa, pd)):
        for a, b in zip(src, dst):
            b.data.copy_(a.data)
    for a, b, c_, d in zip(pa, pb, pc, pd):
        g = torch.from_numpy(rng.standard_normal(a.shape).astype(np.float32)).to(DEV)
        a.grad, b.grad, c_.grad, d.grad = g, g.clone(), g.clone(), g.clone()
    for o in (oa, ob, oc, od):
        o.step()
    for a, b, c_, d in zip(pa, pb, pc, pd):
        assert rel_err(c_.detach().cpu().numpy(), a.detach().cpu().numpy()) < 2e-6
        assert rel_err(d.detach().cpu().numpy(), a.detach().cpu().numpy()) < 2e-6
    assert od.step_count(pd[0]) == 13 and od.step_count(pd[7]) == 12
    sd = ob.state_dict()["state"][0]
    assert sd["step"].dtype == torch.float32 and sd["step"].dim() == 0
    # more tensors than one launch carries (64): several launches, same result
    many_a = [torch.nn.Parameter(torch.full((5,), float(i), device=DEV)) for i in range(70)]
    many_b = [torch.nn.Parameter(p.detach().clone()) for p in many_a]
    oa2, ob2 = torch.optim.Adam(many_a, lr=1e-2, betas=(0.9, 0.99), eps=1e-15), FusedAdam(many_b, lr=1e-2, betas=(0.9, 0.99),
                                                                                       eps=1e-15)
    for _ in range(3):
        for a, b in zip(many_a, many_b):
            a.grad = torch.ones_like(a) * 0.5
            b.grad = torch.ones_like(b) * 0.5
        oa2.step()
        ob2.step()
    for a, b in zip(many_a, many_b):
        assert rel_err(b.detach().cpu().numpy(), a.detach().cpu().numpy()) < 2e-6


@pytest.mark.parametrize("levels,T,mix", [((300, 600), 512, 0), ((40, 700), 16384, 0), ((600,), 300, 1), ((600,), 256, 2)])
def test_node_pass_backward_with_private_tables(levels, T, mix):
    """gngf_node_features_bwd on lattices large enough for the CTA-private table path (shared-memory copies of a level's
    gradient table, k5_encode_bwd.cu) against a float64 torch restatement: table gradients and the adjoint of the
    selected probabilities, all three mix modes."""
    from collision_handling_in_instantngp_b200 import _lib
    F, K = 2, 4
    lat = build_lattice(np.array(levels, dtype=np.int32))
    L, U, S = lat.num_levels, lat.num_nodes, lat.num_level_nodes
    assert max(lat.lwx[l] * lat.lwy[l] for l in range(L)) >= 1 << 18
    g = torch.Generator(device=DEV).manual_seed(T + len(levels))
    tables = [torch.randn((T, F), generator=g, device=DEV) for _ in range(L)]
    tgrads = [torch.zeros((T, F), device=DEV) for _ in range(L)]
    utopv = torch.rand((U, K), generator=g, device=DEV) * 0.5 + 0.01
    utopi = torch.randint(0, T, (U, K), generator=g, device=DEV, dtype=torch.int32)
    dnf = torch.randn((S, F), generator=g, device=DEV)
    dnf[torch.rand(S, generator=g, device=DEV) < 0.3] = 0.0             # untouched level nodes
    dtv = torch.zeros((U, K), device=DEV)
    mode = [_lib.MIX_SOFTMAX, _lib.MIX_WEIGHTED_AVG, _lib.MIX_RAW][mix]
    ops.call("gngf_node_features_bwd", lat, _lib.make_tables(tables), _lib.make_tables(tgrads), T, F, K, mode,
             utopv.data_ptr(), utopi.data_ptr(), dnf.data_ptr(), dtv.data_ptr(), ops._stream())
    torch.cuda.synchronize()
    dtv_ref = torch.zeros((U, K), dtype=torch.float64, device=DEV)
    for l in range(L):
        wx, wy = lat.lwx[l], lat.lwy[l]
        ii = torch.arange(wx * wy, device=DEV)
        cx, cy = lat.lox[l] + ii // wy, lat.loy[l] + ii % wy
        u = (cx - lat.ox) * lat.wy + (cy - lat.oy)
        d = dnf[lat.loff[l]:lat.loff[l] + wx * wy].double()                       # (n, F)
        tv, ti = utopv[u].double(), utopi[u].long()
        if mix == 0:
            w = torch.softmax(tv, -1)
        elif mix == 1:
            w = tv / tv.sum(-1, keepdim=True)
        else:
            w = tv
        rows = tables[l].double()[ti]                                            # (n, K, F)
        ref = torch.zeros((T, F), dtype=torch.float64, device=DEV)
        ref.index_add_(0, ti.reshape(-1), (d[:, None, :] * w[:, :, None]).reshape(-1, F))
        assert float((tgrads[l].double() - ref).abs().max() / ref.abs().max()) < 1e-5, l
        dw = (rows * d[:, None, :]).sum(-1)                                      # (n, K)
        dot = (dw * w).sum(-1, keepdim=True)
        gk = w * (dw - dot) if mix == 0 else ((dw - dot) / tv.sum(-1, keepdim=True) if mix == 1 else dw)
        dtv_ref.index_add_(0, u, gk)
    assert float((dtv.double() - dtv_ref).abs().max() / dtv_ref.abs().max()) < 1e-5
