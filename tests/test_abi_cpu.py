"""CPU: the C-ABI library is built, loads, and exports every symbol include/gngf.h declares."""
import os
import re

import collision_handling_in_instantngp_b200 as pkg
from collision_handling_in_instantngp_b200 import _lib
from collision_handling_in_instantngp_b200.lattice import build_lattice, level_resolutions

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gngf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gngf_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = pkg.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in gngf.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype"
    assert sorted(_lib.SIGNATURES) == names
    assert lib.gngf_abi_version() == 1
    assert lib.gngf_strerror(0) == b"ok"
    assert lib.gngf_strerror(-1) == b"invalid argument"
    assert lib.gngf_launch_count() == 0


def test_library_has_sm100a_code_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", pkg.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_lattice_boxes():
    n_ls = level_resolutions(8, 32, 4)
    assert list(n_ls) == [8, 12, 20, 32]
    lat = build_lattice(n_ls)
    assert (lat.ox, lat.oy, lat.wx, lat.wy) == (0, 0, 34, 34)
    assert [lat.lwx[l] for l in range(4)] == [10, 14, 22, 34]
    assert lat.num_level_nodes == 100 + 196 + 484 + 1156
    # the strawberry image: x in [0,1], y in [0, 338/507]
    lat = build_lattice(n_ls, (0.0, 0.0), (1.0, 338 / 507))
    assert lat.num_nodes == 34 * 23     # the 782 distinct HPD inputs of SURVEY.md section 7
    lat = build_lattice(level_resolutions(16, 8192, 16))
    assert lat.wx == 8193 and lat.n[15] == 8191
    # negative coordinates (BatchNorm'd inputs): boxes follow the data
    lat = build_lattice(n_ls, (-1.5, -0.25), (2.0, 0.5))
    assert lat.lox[0] == -12 and lat.loy[0] == -2 and lat.lwx[0] == 16 - (-12) + 2


def test_host_side_argument_validation_needs_no_gpu():
    """Entry points reject bad arguments before touching the device; FusedAdam refuses CPU parameters (no CPU path)."""
    import ctypes

    import pytest
    import torch

    from collision_handling_in_instantngp_b200.optim import FusedAdam
    lib = pkg.load()
    assert lib.gngf_mlp3_tc_supported(8, 64, 64, 3) == 1 and lib.gngf_mlp3_tc_supported(65, 64, 64, 3) == 0
    assert lib.gngf_mlp3_tc_supported(8, 32, 64, 3) == 0
    arr = (_lib.AdamTensor * 1)()
    assert lib.gngf_adam_step(arr, 65, 0.9, 0.99, 1e-15, None, None) == -1          # too many tensors / no ticket
    assert lib.gngf_peer_allreduce(None, None, 0, 2, None, None, 4, 4, 1, 1.0, None, None) == -1
    assert ctypes.sizeof(_lib.AdamTensor) == 56
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    opt = FusedAdam([p], lr=1e-3)
    with pytest.raises(_lib.GngfError):
        opt.step()
    with pytest.raises(ValueError):
        FusedAdam([p], lr=-1.0)


def test_streaming_backward_workspace_and_stats_entry_points():
    """The workspace of the streaming backward holds five per-node vectors (a, -max log2e, column offset, sign, and the
    screening product's error bound), the dlogits of the K selected slots and the device scalars; the statistics entry
    point rejects a null buffer before it touches the device."""
    lib = pkg.load()
    for U, K in ((1, 1), (130, 4), (173400, 4), (29972809, 8)):
        U4 = (U + 3) // 4 * 4
        assert lib.gngf_hpd_stream_bwd_workspace_floats(U, K) == 5 * U4 + U * K + 16
    assert lib.gngf_hpd_stream_bwd_stats(None, 0) == -1


def test_active_node_and_split_loss_entry_points_validate_on_the_host():
    """k11_active_nodes.cu / gngf_loss_parts / the *_enc and *_nodes variants: sizes and bad arguments are answered on
    the host, before any launch."""
    lib = pkg.load()
    lat = build_lattice(level_resolutions(16, 8192, 16))
    U = lat.num_nodes
    assert U == 8193 * 8193
    assert lib.gngf_active_nodes_bitmap_words(U) == (U + 31) // 32
    assert lib.gngf_active_nodes_bitmap_words(0) == 0 and lib.gngf_active_nodes_bitmap_words(33) == 2
    words = (U + 31) // 32
    assert lib.gngf_active_nodes_chunks(U) == (words + 1023) // 1024            # one block per 1024 words = 32768 nodes
    assert lib.gngf_active_nodes_chunks(0) == 0
    assert lib.gngf_lattice_mark_nodes(None, 0, lat, None, None) == -1          # no bitmap
    assert lib.gngf_lattice_mark_nodes(None, -1, lat, 16, None) == -1
    assert lib.gngf_lattice_mark_nodes(None, 0, lat, 16, None) == 0             # empty batch: nothing to launch
    assert lib.gngf_compact_nodes(None, U, None, None, 0, None, None) == -1
    assert lib.gngf_compact_nodes(16, 1 << 31, 16, 16, 1, 16, None) == -1       # node ids are int32
    assert lib.gngf_compact_nodes(20, 64, 16, 16, 1, 16, None) == -1            # bitmap must be 16-byte aligned
    assert lib.gngf_scatter_node_rows(None, 0, None, 4, None, None) == 0        # no rows
    assert lib.gngf_scatter_node_rows(None, 5, None, 4, None, None) == -1
    assert lib.gngf_scatter_node_rows(16, 5, 16, 0, 16, None) == -1
    assert lib.gngf_hpd_first_layer_fwd_nodes(lat, 16, U + 1, None, None, 32, 1, None, None) == -1   # more rows than nodes
    assert lib.gngf_hpd_first_layer_fwd_nodes(lat, 16, 0, None, None, 32, 1, None, None) == 0        # empty list
    assert lib.gngf_hpd_first_layer_bwd_nodes(lat, 16, 0, None, 32, None, None, None) == 0
    # loss halves: parts must name a half, each half needs its own buffers
    args = (None, None, 0, None, 4, 256, 1.0, -2.0, 1.0, 1.0, 1.0, None)
    assert lib.gngf_loss_parts(*args, 16, None, None, 0, None) == -1
    assert lib.gngf_loss_parts(*args, None, None, None, 3, None) == -1
    assert lib.gngf_loss_parts(*args, 16, None, None, 1, None) == -1            # MSE half without rgb / target / d_rgb
    assert lib.gngf_loss_parts(*args, 16, None, None, 2, None) == -1            # divergence half without column sums
    # the streaming backward on a node list accepts fewer rows than the box; the plain entry point only without a
    # column-sum adjoint (rows whose adjoints arrive row-indexed: the node-parallel owner's call)
    small = build_lattice(level_resolutions(8, 32, 4))
    a = (16,) * 7          # h_planes, h_scale, w_planes, w_scale, h, w, bias
    common = (16, 16, 16, None, None, 16, 16, 1, 16, 16, 16, 16, None)
    with_colsum = (16, 16, 16, 16, 16, 16, 16, 1, 16, 16, 16, 16, None)
    assert lib.gngf_hpd_stream_bwd(small, *a, 100, 256, 128, 4, *with_colsum) == -2               # U != box
    # node-parallel helpers
    assert lib.gngf_bitmap_or(None, 2, 8, None, None) == -1
    assert lib.gngf_bitmap_or(16, 2, 6, 16, None) == -1                         # whole 16-byte groups of words only
    assert lib.gngf_bitmap_or(16, 0, 8, 16, None) == -1
    assert lib.gngf_bitmap_or(16, 2, 0, 16, None) == 0
    assert lib.gngf_gather_node_adjoints(small, None, 0, 4, None, None, None, None, None) == 0    # empty list
    assert lib.gngf_gather_node_adjoints(small, None, 5, 4, None, None, None, 16, None) == -1     # no adjoints
    assert lib.gngf_gather_node_adjoints(small, None, 5, 4, 16, None, 16, 16, None) == -1         # gcol_k without cnt
    assert lib.gngf_gather_node_adjoints(small, None, 5, 0, 16, None, None, 16, None) == -1
    assert lib.gngf_split_f16x2(16, -1, 16, 16, None) == -1 and lib.gngf_split_f16x2(16, 8, 16, None, None) == -1
    assert lib.gngf_split_f16x2(20, 8, 16, 16, None) == -1                      # source must be 16-byte aligned
    assert lib.gngf_tc_gemm_set_formats(2, 0) == -1 and lib.gngf_tc_gemm_set_formats(1, 1) == 0
    assert lib.gngf_peer_allreduce_set_timeout_ms(0) == -1 and lib.gngf_peer_allreduce_set_timeout_ms(30000) == 0
    assert lib.gngf_hpd_stream_bwd_nodes(small, 16, *a, small.num_nodes + 1, 256, 128, 4, *common) == -2
