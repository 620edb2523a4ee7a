"""PyTorch-side plumbing of the GNGF hot path: output allocation, stream hand-off and autograd wiring around
the C ABI of libgngf_sm100.so.  PyTorch owns device memory and streams; all arithmetic is in the CUDA library.

The path (reference models.py:394-484 and its autograd) is organised around the *lattice* (see lattice.py):

  forward   HPD MLP + softmax + top-k on the U lattice nodes            (K2, K3)
            mixed feature per level node                                 (K4 node pass)
            per point: corners, bilinear weights, 4 gathers per level    (K1+K4 point pass) -> enc (P, L*F)
            decoder MLP                                                  (K6)                -> rgb
            node multiplicities -> column sums of the (virtual) probs    (what Loss consumes, utils.py:138)
  backward  decoder -> d enc -> per-level-node sums -> table gradients + top-k adjoint (K5a)
            softmax / top-k / column-sum adjoint per node -> dlogits     (K5b) -> HPD layer gradients (K5c)
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import torch

from . import _lib
from ._lib import (ACT_LEAKY_RELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, MIX_RAW, MIX_SOFTMAX, MIX_WEIGHTED_AVG, GngfError,
                   Lattice, call, make_tables)

MAX_DENSE_LOGIT_BYTES = 96 << 30   # (U, T) fp32 logits are materialised by this path

# data parallelism (dp.enable_gradient_allreduce):
#   GRAD_REDUCE_HOOK(flat_params)      in-place mean over ranks of the flat gradient buffer, at the end of GNGFPath.backward
#   COLSUM_REDUCE_HOOK(colsum)  in-place sum over ranks of the (L, N) column sums, right behind the kernel that produces
#                               them -- on the side stream, so the exchange overlaps the decoder forward; returns the
#                               world size (the adjoint of the local column sums is world * the adjoint of the sum,
#                               because every rank evaluates the same function of it)
#   NODE_SHARD                  dp.NodeSharding: the streaming HPD path evaluates 1/world of the touched lattice nodes
#                               per rank (selections all-gathered, adjoints reduce-scattered) instead of all of them
GRAD_REDUCE_HOOK = None
COLSUM_REDUCE_HOOK = None
NODE_SHARD = None


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# ---- intra-step concurrency -------------------------------------------------------------------------------------
# At the published batch size every kernel of the step is a few tens of microseconds and fills a fraction of the
# chip, so independent kernels are forked onto side streams (the API-output gather of idx_topk, the column sums the
# loss needs, the per-layer weight-gradient products of the HPD) and joined before their results are consumed.
# Under CUDA-graph capture the forks become parallel branches of the graph.  Buffers are always allocated on the
# main stream (the caching allocator ties a block to the stream it was allocated on); only launches move.
CONCURRENT = True
# trainer.GraphedTrainer: the forward returns WITHOUT joining the side stream that produces the column sums (and, under
# data parallelism, all-reduces them); the caller evaluates the divergence half of the loss on that stream
# (state.colsum_fork.side) and the backward joins it after the encoding's point pass, right before the column-sum
# adjoint is consumed.  The column-sum branch then overlaps the decoder forward AND backward instead of gating the loss.
DEFER_COLSUM_JOIN = False
# The API output idx_topk (P,L,4,K) int64 (models.py:476-484) is a pure gather that nothing inside a training step reads
# (8.6 GB per step at BASELINE.json configs[3]); callers that never look at it -- trainer.GraphedTrainer -- switch it off
WANT_IDX_TOPK = True
_SIDE_STREAMS = {}


def _side_stream(dev, i: int) -> torch.cuda.Stream:
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), i)
    s = _SIDE_STREAMS.get(key)
    if s is None:
        s = _SIDE_STREAMS[key] = torch.cuda.Stream(device=dev)
    return s


class _Fork:
    """with _Fork(dev, i): ...launches...   -- runs the block on side stream i after everything queued so far on the
    current stream; `.join()` makes the current stream wait for it.  With CONCURRENT off it is a no-op."""

    def __init__(self, dev, i: int):
        self.on = CONCURRENT
        if self.on:
            self.main = torch.cuda.current_stream(dev)
            self.side = _side_stream(dev, i)
            self.ctx = torch.cuda.stream(self.side)

    def __enter__(self):
        if self.on:
            self.side.wait_stream(self.main)
            self.ctx.__enter__()
        return self

    def __exit__(self, *exc):
        if self.on:
            self.ctx.__exit__(*exc)
        return False

    def join(self):
        if self.on:
            self.main.wait_stream(self.side)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise GngfError(f"{what} must be a CUDA tensor (got {t.device}); this package has no CPU path")


def mix_mode_of(flag) -> int:
    """params.should_softmax_topk_features: True / False / None (models.py:212-217)."""
    if flag is None:
        return MIX_RAW
    return MIX_SOFTMAX if flag else MIX_WEIGHTED_AVG


# ----------------------------------------------------------------------------------------------------------
# thin wrappers (one C call each)
# ----------------------------------------------------------------------------------------------------------
def corners_fwd(x: torch.Tensor, lat: Lattice):
    """_scale_to_grid (models.py:486-502): scaled (P,2,L,1), grid (P,2,L,4), fp32."""
    _require_cuda(x, "x")
    x = _f32c(x)
    P, L = x.shape[0], lat.num_levels
    scaled = torch.empty((P, 2, L, 1), dtype=torch.float32, device=x.device)
    grid = torch.empty((P, 2, L, 4), dtype=torch.float32, device=x.device)
    call("gngf_corners_fwd", x.data_ptr(), P, lat, scaled.data_ptr(), grid.data_ptr(), _stream())
    return scaled, grid


def fast_hash_fwd(x: torch.Tensor, lat: Lattice, table_size: int) -> torch.Tensor:
    """_fast_hash (models.py:504-528) of the corners of x: (P,L,4) int64."""
    _require_cuda(x, "x")
    x = _f32c(x)
    idx = torch.empty((x.shape[0], lat.num_levels, 4), dtype=torch.int64, device=x.device)
    call("gngf_fast_hash_fwd", x.data_ptr(), x.shape[0], lat, int(table_size), idx.data_ptr(), _stream())
    return idx


def linear_fwd(x, w, b, act) -> torch.Tensor:
    M, K = x.shape
    N = w.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    call("gngf_linear_fwd", x.data_ptr(), w.data_ptr(), _ptr(b), M, N, K, act, y.data_ptr(), _stream())
    return y


def linear_bwd(dz, x, w, act_prev, dx_needed, dw, db) -> Optional[torch.Tensor]:
    M, N = dz.shape
    K = w.shape[1]
    dx = torch.empty((M, K), dtype=torch.float32, device=dz.device) if dx_needed else None
    call("gngf_linear_bwd", dz.data_ptr(), x.data_ptr(), w.data_ptr(), M, N, K, act_prev, _ptr(dx), _ptr(dw), _ptr(db),
         _stream())
    return dx


def split_bf16x3(t: torch.Tensor) -> torch.Tensor:
    """(3, *t.shape) bf16 planes with hi + mid + lo == t to ~2^-24 relative."""
    t = _f32c(t)
    planes = torch.empty((3, *t.shape), dtype=torch.bfloat16, device=t.device)
    call("gngf_split_bf16x3", t.data_ptr(), t.numel(), planes.data_ptr(), _stream())
    return planes


class Planes16:
    """Two fp16 planes of a power-of-two-scaled fp32 tensor (gngf_split_f16x2): the operands of the streaming HPD
    kernels.  `data` (2, *shape) float16; `scale` 2 device floats (scale[0] = 2^-s, undone in the consumers' epilogues)."""
    __slots__ = ("data", "scale")

    def __init__(self, data, scale):
        self.data, self.scale = data, scale


def split_f16x2(t: torch.Tensor) -> Planes16:
    """x 2^s = hi + mid in fp16 with the scale chosen on the device from max|x| (no host synchronisation)."""
    t = _f32c(t)
    planes = torch.empty((2, *t.shape), dtype=torch.float16, device=t.device)
    scale = torch.empty(2, dtype=torch.float32, device=t.device)
    call("gngf_split_f16x2", t.data_ptr(), t.numel(), planes.data_ptr(), scale.data_ptr(), _stream())
    return Planes16(planes, scale)


def tc_linear_fwd(x, w, b, act, x_planes=None, w_planes=None) -> torch.Tensor:
    """y = act(x w^T + b) on the tensor cores (tcgen05, split-bf16 operands, fp32 accumulate in TMEM)."""
    M, K = x.shape
    N = w.shape[0]
    xp = split_bf16x3(x) if x_planes is None else x_planes
    wp = split_bf16x3(w) if w_planes is None else w_planes
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    call("gngf_tc_gemm_bf16x3", xp.data_ptr(), wp.data_ptr(), _ptr(b), M, N, K, act, 0, 1, y.data_ptr(), _stream())
    return y


def split_bf16x3_t(t: torch.Tensor) -> torch.Tensor:
    """bf16 planes of t^T: (3, cols, ld) with ld = rows rounded up to 8 (16-byte row pitch), zero padded."""
    t = _f32c(t)
    rows, cols = t.shape
    ld = (rows + 7) & ~7
    planes = torch.empty((3, cols, ld), dtype=torch.bfloat16, device=t.device)
    call("gngf_split_bf16x3_t", t.data_ptr(), rows, cols, ld, planes.data_ptr(), _stream())
    return planes


def tc_gemm_planes(a_planes, b_planes, out, accumulate=False, k_splits=1) -> None:
    """out (M,N) (+)= A B^T for bf16x3 planes a (3,M,K), b (3,N,K) on the tensor cores.  k_splits > 1 sums K
    ranges with atomics: `out` must then be zero-initialised."""
    M, K = a_planes.shape[1], a_planes.shape[2]
    N = b_planes.shape[1]
    call("gngf_tc_gemm_bf16x3", a_planes.data_ptr(), b_planes.data_ptr(), None, M, N, K, ACT_NONE, int(accumulate),
         int(k_splits), out.data_ptr(), _stream())


STREAM_FWD_REFINED = True         # streaming forward as a two-plane pass + fp32 refinement of 8 candidates (K <= 4)


def hpd_stream_fwd(h, w, b, k, h_planes=None, w_planes=None, alloc_rows=None):
    """Fused K2+K3: (utopv, utopi int32, row_max, row_sum) of softmax(h w^T + b) without the (U,T) logits.
    K <= 4: half the tensor-core work -- a two-plane streaming pass (fp16 planes of the scaled operands, Planes16) keeps
    8 candidates per row, whose logits are then re-evaluated in fp32 and re-ranked (gngf_hpd_stream_fwd_refined);
    otherwise three bf16 planes / six products.  h_planes / w_planes: Planes16 (what the backward uses as well)."""
    U, Kd = h.shape
    T = w.shape[0]
    refined = STREAM_FWD_REFINED and k <= 4
    dev = h.device
    # alloc_rows >= U: the selections are written into the head of larger (zero-tailed) buffers -- the equal-sized
    # per-rank blocks of dp.NodeSharding.all_gather_rows
    if alloc_rows is None:
        utopv = torch.empty((U, k), dtype=torch.float32, device=dev)
        utopi = torch.empty((U, k), dtype=torch.int32, device=dev)
    else:
        utopv = torch.zeros((alloc_rows, k), dtype=torch.float32, device=dev)
        utopi = torch.zeros((alloc_rows, k), dtype=torch.int32, device=dev)
    if U == 0:
        return utopv, utopi, torch.empty(0, dtype=torch.float32, device=dev), torch.empty(0, dtype=torch.float32, device=dev)
    row_max = torch.empty(U, dtype=torch.float32, device=dev)
    row_sum = torch.empty(U, dtype=torch.float32, device=dev)
    if refined:
        h, w = _f32c(h), _f32c(w)
        hp = split_f16x2(h) if h_planes is None else h_planes
        wp = split_f16x2(w) if w_planes is None else w_planes
        work = torch.empty(_lib.load().gngf_hpd_stream_refined_workspace_floats(U, T), dtype=torch.float32, device=dev)
        call("gngf_hpd_stream_fwd_refined", hp.data.data_ptr(), hp.scale.data_ptr(), wp.data.data_ptr(),
             wp.scale.data_ptr(), h.data_ptr(), w.data_ptr(), b.data_ptr(), U, T, Kd, k, utopv.data_ptr(),
             utopi.data_ptr(), row_max.data_ptr(), row_sum.data_ptr(), work.data_ptr(), _stream())
        return utopv, utopi, row_max, row_sum
    hp, wp = split_bf16x3(h), split_bf16x3(w)      # (K > 4: three bf16 planes, six products)
    work = torch.empty(_lib.load().gngf_hpd_stream_workspace_floats(U, T, k), dtype=torch.float32, device=dev)
    call("gngf_hpd_stream_fwd", hp.data_ptr(), wp.data_ptr(), b.data_ptr(), U, T, Kd, k, utopv.data_ptr(),
         utopi.data_ptr(), row_max.data_ptr(), row_sum.data_ptr(), work.data_ptr(), _stream())
    return utopv, utopi, row_max, row_sum


def hpd_stream_bwd(lat: Lattice, h, w, b, h_planes, w_planes, utopv, utopi, dtv, cnt, gcol_k, row_max, row_sum,
                   dw, db, act_prev=ACT_RELU, node_ids=None) -> torch.Tensor:
    """Fused K5c (k2_hpd_tc_bwd.cu): returns dh (U,Kd) = (dlogits w) .* act'(h); accumulates dlogits^T h into dw and
    colsum(dlogits) into db, with dlogits = -<g,p_top> p + scatter(p_k g_k) never materialised."""
    U, kd = h.shape
    T, K = w.shape[0], utopv.shape[1]
    dh = torch.zeros((U, kd), dtype=torch.float32, device=h.device)
    work = torch.empty(_lib.load().gngf_hpd_stream_bwd_workspace_floats(U, K), dtype=torch.float32, device=h.device)
    if node_ids is None:
        call("gngf_hpd_stream_bwd", lat, h_planes.data.data_ptr(), h_planes.scale.data_ptr(), w_planes.data.data_ptr(),
             w_planes.scale.data_ptr(), h.data_ptr(), w.data_ptr(), b.data_ptr(), U, T, kd, K, utopv.data_ptr(), utopi.data_ptr(), dtv.data_ptr(), _ptr(cnt), _ptr(gcol_k),
             row_max.data_ptr(), row_sum.data_ptr(), act_prev, dh.data_ptr(), dw.data_ptr(), db.data_ptr(),
             work.data_ptr(), _stream())
    else:   # rows = active nodes; dtv and cnt stay indexed by the lattice node
        call("gngf_hpd_stream_bwd_nodes", lat, node_ids.data_ptr(), h_planes.data.data_ptr(),
             h_planes.scale.data_ptr(), w_planes.data.data_ptr(), w_planes.scale.data_ptr(), h.data_ptr(), w.data_ptr(),
             b.data_ptr(), U, T, kd, K, utopv.data_ptr(), utopi.data_ptr(), dtv.data_ptr(),
             _ptr(cnt), _ptr(gcol_k), row_max.data_ptr(), row_sum.data_ptr(), act_prev, dh.data_ptr(), dw.data_ptr(),
             db.data_ptr(), work.data_ptr(), _stream())
    return dh


def softmax_topk_fwd(logits: torch.Tensor, k: int, inplace: bool = False, want_probs: bool = True):
    """probs = nan_to_num(softmax(logits)), (topv, topi) = topk(probs, k) with ties -> lower index."""
    _require_cuda(logits, "logits")
    logits = _f32c(logits)
    R, T = logits.shape
    probs = (logits if inplace else torch.empty_like(logits)) if want_probs else None
    topv = torch.empty((R, k), dtype=torch.float32, device=logits.device)
    topi = torch.empty((R, k), dtype=torch.int32, device=logits.device)
    call("gngf_softmax_topk_fwd", logits.data_ptr(), R, T, k, _ptr(probs), topv.data_ptr(), topi.data_ptr(), None, None,
         _stream())
    return probs, topv, topi


def topk_fwd(values: torch.Tensor, k: int):
    _require_cuda(values, "input")
    values = _f32c(values)
    shape = values.shape
    v2 = values.reshape(-1, shape[-1])
    topv = torch.empty((v2.shape[0], k), dtype=torch.float32, device=values.device)
    topi = torch.empty((v2.shape[0], k), dtype=torch.int64, device=values.device)
    call("gngf_topk_fwd", v2.data_ptr(), v2.shape[0], v2.shape[1], k, topv.data_ptr(), topi.data_ptr(), _stream())
    return topv.reshape(*shape[:-1], k), topi.reshape(*shape[:-1], k)


def topk_bwd(grad_values: torch.Tensor, topi: torch.Tensor, T: int) -> torch.Tensor:
    gv = _f32c(grad_values)
    shape = gv.shape
    g2 = gv.reshape(-1, shape[-1])
    i2 = topi.reshape(-1, shape[-1]).contiguous()
    gin = torch.empty((g2.shape[0], T), dtype=torch.float32, device=gv.device)
    call("gngf_topk_bwd", g2.data_ptr(), i2.data_ptr(), g2.shape[0], T, shape[-1], gin.data_ptr(), _stream())
    return gin.reshape(*shape[:-1], T)


def count_distinct(indices: torch.Tensor, value_range: int):
    """Distinct integer values in [0, value_range) per (column, level) of an index tensor (P,L,V,C), float32 or
    int64 -> (uniq (C,L) int32, outliers flag tensor).  The kernel behind calc_hash_collisions."""
    _require_cuda(indices, "indices")
    if indices.dtype not in (torch.float32, torch.int64):
        indices = indices.float() if indices.is_floating_point() else indices.long()
    indices = indices.contiguous()
    P, L, V, C = indices.shape
    dev = indices.device
    words = _lib.load().gngf_count_distinct_workspace_words(L, C, int(value_range))
    bitmap = torch.empty(words, dtype=torch.int32, device=dev)
    uniq = torch.empty((C, L), dtype=torch.int32, device=dev)
    outliers = torch.empty(1, dtype=torch.int32, device=dev)
    name = "gngf_count_distinct_f32" if indices.dtype == torch.float32 else "gngf_count_distinct_i64"
    call(name, indices.data_ptr(), P, L, V, C, int(value_range), bitmap.data_ptr(), uniq.data_ptr(),
         outliers.data_ptr(), _stream())
    return uniq, outliers


def counts_per_level(grid: torch.Tensor, lat: Lattice, hashed: torch.Tensor, table_size: int):
    """_calc_counts_per_level (models.py:530-566) on the GPU: grid (P,2,L,4) fp32 corners, hashed (P,L,4) int64 slots
    (any view with a uniform element stride, e.g. idx_topk[..., 0]) -> (hist (L,T) int32, outliers flag tensor)."""
    _require_cuda(grid, "grid")
    grid = _f32c(grid)
    P, L = grid.shape[0], lat.num_levels
    if hashed.dtype != torch.int64:
        hashed = hashed.long()
    # (P,L,4) with strides (L*4*s, 4*s, s): a column of a contiguous (P,L,4,K) tensor, or a contiguous (P,L,4) tensor
    s = hashed.stride(2) if P > 0 else 1
    if P > 0 and (hashed.stride(1) != 4 * s or hashed.stride(0) != L * 4 * s or s < 1):
        hashed = hashed.contiguous()
        s = 1
    dev = grid.device
    first = torch.empty(lat.num_level_nodes, dtype=torch.int32, device=dev)
    hist = torch.empty((L, int(table_size)), dtype=torch.int32, device=dev)
    outliers = torch.empty(1, dtype=torch.int32, device=dev)
    call("gngf_counts_per_level", grid.data_ptr(), P, lat, hashed.data_ptr(), s, int(table_size), first.data_ptr(),
         hist.data_ptr(), outliers.data_ptr(), _stream())
    return hist, outliers


def gather_rows(x, lat: Lattice, uvals: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(P,L,4,N) rows of a per-node array; int32 input gives the int64 API dtype."""
    P, N = x.shape[0], uvals.shape[1]
    if uvals.dtype == torch.int32:
        if out is None:
            out = torch.empty((P, lat.num_levels, 4, N), dtype=torch.int64, device=x.device)
        call("gngf_lattice_gather_rows_i64", x.data_ptr(), P, lat, uvals.data_ptr(), N, out.data_ptr(), _stream())
    else:
        if out is None:
            out = torch.empty((P, lat.num_levels, 4, N), dtype=torch.float32, device=x.device)
        call("gngf_lattice_gather_rows", x.data_ptr(), P, lat, uvals.data_ptr(), N, out.data_ptr(), _stream())
    return out


def stream_bwd_stats(reset=False):
    """[dh-pass tiles, of which needed all three logit products, of which issued the second product, dW3-pass tiles, ...,
    ...] since the last reset (gngf.h: gngf_hpd_stream_bwd_stats); synchronises the device.  bench.py's executed-FLOP
    accounting."""
    import ctypes
    out = (ctypes.c_uint64 * 6)()
    call("gngf_hpd_stream_bwd_stats", ctypes.cast(out, ctypes.c_void_p), 1 if reset else 0)
    return [int(v) for v in out]


def active_nodes(x: torch.Tensor, lat: Lattice, shard=None) -> torch.Tensor:
    """Ascending int32 ids of the lattice nodes that the corners of x touch (k11_active_nodes.cu).  One
    device->host read of the count (the list sizes every launch of the HPD chain).  With `shard` (dp.NodeSharding) the
    per-rank bitmaps are OR-ed over the ranks first: every rank gets the same list -- the nodes ANY rank touches."""
    _require_cuda(x, "x")
    x = _f32c(x)
    lib = _lib.load()
    U, P = lat.num_nodes, x.shape[0]
    dev = x.device
    words = int(lib.gngf_active_nodes_bitmap_words(U))
    bitmap = torch.zeros((words + 3) & ~3, dtype=torch.int32, device=dev)
    call("gngf_lattice_mark_nodes", x.data_ptr(), P, lat, bitmap.data_ptr(), _stream())
    if shard is not None:
        bitmap = shard.union_bitmap(bitmap)
    chunk_offsets = torch.empty(int(lib.gngf_active_nodes_chunks(U)), dtype=torch.int32, device=dev)
    capacity = max(1, min(U, P * lat.num_levels * 4 * (1 if shard is None else shard.world)))
    ids = torch.empty(capacity, dtype=torch.int32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    call("gngf_compact_nodes", bitmap.data_ptr(), U, chunk_offsets.data_ptr(), ids.data_ptr(), capacity,
         count.data_ptr(), _stream())
    return ids[:int(count.item())]


def scatter_node_rows(node_ids: torch.Tensor, src: torch.Tensor, dst: torch.Tensor) -> None:
    """dst[node_ids[r], :] = src[r, :] for 32-bit element types."""
    assert src.element_size() == 4 and dst.dtype == src.dtype and src.is_contiguous() and dst.is_contiguous()
    call("gngf_scatter_node_rows", node_ids.data_ptr(), node_ids.shape[0], src.data_ptr(), src.shape[1], dst.data_ptr(),
         _stream())


# ----------------------------------------------------------------------------------------------------------
# the fused forward / backward
# ----------------------------------------------------------------------------------------------------------
@dataclass
class PathConfig:
    """Static description of one GeneralNeuralGaugeFields instance."""
    table_size: int
    feature_dim: int
    topk_k: int
    n_hpd: int                     # number of HPD linear layers
    n_mlp: int                     # number of decoder linear layers
    topk_only: bool = False        # should_keep_topk_only
    mix_mode: int = MIX_SOFTMAX
    leaky: bool = False            # params.should_leaky_relu
    use_hash: bool = False         # params.should_use_hash_function
    hpd_trainable: bool = True
    drop_topk_adjoint: bool = False  # params.should_inplace_scatter is None: DifferentiableTopk.backward returns zeros
                                     # (models.py:30-31 discards the scatter result) -> nothing reaches the HPD through
                                     # the selected probabilities


@dataclass
class ForwardState:
    """What one forward leaves behind for backward and for the lazily materialised outputs."""
    x: torch.Tensor
    lat: Lattice
    cfg: PathConfig
    hpd_acts: List[torch.Tensor] = field(default_factory=list)   # outputs of HPD layers 0..n-2, (U, width)
    uprobs: Optional[torch.Tensor] = None                        # (U,T); None on the streaming path
    row_max: Optional[torch.Tensor] = None                       # (U,) softmax statistics (streaming path)
    row_sum: Optional[torch.Tensor] = None
    w_planes: Optional["Planes16"] = None                        # (2,T,Kd) fp16 planes + scale of the output layer
    h_planes: Optional["Planes16"] = None                        # (2,U,Kd) fp16 planes + scale of its input (streaming path)
    hpd_small: bool = False                                      # fused small-lattice HPD kernels were used
    nfeat: Optional[torch.Tensor] = None                         # (S,F) written by the HPD kernel itself (small lattices)
    utopv: Optional[torch.Tensor] = None                         # (U,K)
    utopi: Optional[torch.Tensor] = None                         # (U,K) int32
    shard: Optional[tuple] = None                                # (dp.NodeSharding, listed nodes, r0, r1, chunk): this rank
    node_ids_all: Optional[torch.Tensor] = None                  #   evaluated rows r0..r1 of the agreed node list
    node_ids: Optional[torch.Tensor] = None                      # (Ua,) int32 active nodes: the HPD chain (hpd_acts,
    utopv_rows: Optional[torch.Tensor] = None                    #   h_planes, row_max / row_sum, utopv_rows / utopi_rows)
    utopi_rows: Optional[torch.Tensor] = None                    #   has one row per ACTIVE node; utopv / utopi stay (U,K)
    cnt: Optional[torch.Tensor] = None                           # (S,) int32
    mlp_acts: List[torch.Tensor] = field(default_factory=list)   # enc, a1, ..., rgb  (enc, rgb when fused)
    mlp_fused: bool = False
    mlp_tc: bool = False                                         # decoder ran on the tensor cores (k6_mlp_tc.cu)
    mlp_masks: Optional[torch.Tensor] = None                     # (P,4) int32 ReLU pattern of the hidden layers
    idx_topk: Optional[torch.Tensor] = None                      # (P,L,4,K) int64, the API output (models.py:476-484)
    colsum_world: int = 1                                        # > 1: the column sums were summed over that many ranks
    colsum_fork: Optional["_Fork"] = None                        # un-joined column-sum branch (DEFER_COLSUM_JOIN)
    err_flag: Optional[torch.Tensor] = None


def _mlp3_supported(mlp_w) -> bool:
    """True when the decoder has the reference's shape IN -> 64 -> 64 -> OUT (fused kernels of k6_mlp.cu)."""
    if len(mlp_w) != 3:
        return False
    return bool(_lib.load().gngf_mlp3_supported(mlp_w[0].shape[1], mlp_w[0].shape[0], mlp_w[1].shape[0],
                                                mlp_w[2].shape[0])) and mlp_w[1].shape[1] == mlp_w[0].shape[0]


# streaming (never-materialised) HPD output layer: used in top-k-only mode once U*T is large; tests can force it
STREAM_MIN_ELEMENTS = 1 << 26
FORCE_STREAMING = None            # None: by size; True / False: override
TC_MIN_ELEMENTS = 1 << 22         # dense logits come from the tensor-core GEMM above this size
BWD_CHUNK_BYTES = 4 << 30         # logits recomputed per chunk of rows in the (un-fused) streaming backward
STREAM_BWD_FUSED = True           # streaming backward as the fused tcgen05 kernels of k2_hpd_tc_bwd.cu (False: the
                                  # chunked recompute through the plain GEMM, kept as a cross-check for the tests)


# active nodes (k11_active_nodes.cu): on large lattices the streaming HPD path is evaluated only on the nodes the batch
# touches (BASELINE.json configs[3]: 2^22 points touch ~40 % of the 67 M nodes of the 8192^2 box; a data-parallel rank
# with 1/8 of the batch ~12 %).  Costs one device->host read of the node count per forward.
ACTIVE_NODES_MIN = 1 << 20        # lattices with fewer nodes are evaluated densely
ACTIVE_NODES_MAX_FRACTION = 0.75  # ... and so are batches that touch more than this share of the box
FORCE_ACTIVE_NODES = None         # None: by size; True / False: override (tests)


def _active_nodes_wanted(U: int) -> bool:
    if not STREAM_BWD_FUSED:
        return False
    if torch.cuda.is_current_stream_capturing():
        return False        # the node count is read back to size the launches: not possible inside a CUDA-graph capture
    if FORCE_ACTIVE_NODES is not None:
        return bool(FORCE_ACTIVE_NODES)
    return U >= ACTIVE_NODES_MIN


def _streaming_ok(cfg, U, T, k, kd) -> bool:
    if not cfg.topk_only or k > 8 or kd > 128 or kd % 8 != 0:
        return False
    if FORCE_STREAMING is not None:
        return bool(FORCE_STREAMING)
    return U * T >= STREAM_MIN_ELEMENTS


MLP_TENSOR_CORES = True           # decoder through k6_mlp_tc.cu (tcgen05); False: the fp32 CUDA-core kernels of k6_mlp.cu

SMALL_LATTICE_MAX_NODES = 8192    # fused per-node HPD kernels (k2_hpd_small.cu) below this many nodes
SMALL_FUSE_NODE_PASSES = True     # ... which then also run the encoding's per-level-node passes (two launches fewer on the
                                  # critical path of the step); False: separate node_features_fwd / _bwd launches


def _hpd_small_ok(hpd_w, U, k) -> bool:
    if U > SMALL_LATTICE_MAX_NODES or len(hpd_w) < 2:
        return False
    widths = [hpd_w[0].shape[1]] + [w.shape[0] for w in hpd_w]
    return bool(_lib.load().gngf_hpd_small_supported(len(hpd_w), _lib.int_array(widths), k))


def hpd_forward_nodes(lat: Lattice, hpd_w, hpd_b, k: int, device, cfg, state, tables=None):
    """HashProbDistribution.forward (models.py:90-123) on every lattice node.  Fills state.hpd_acts and
    state.utopv / utopi (U,K); state.uprobs (U,T) on the dense path, state.row_max / row_sum (U) on the
    streaming path (tcgen05 GEMM fused with online softmax + running top-k, logits never written)."""
    U = lat.num_nodes
    T = hpd_w[-1].shape[0]
    n = len(hpd_w)
    if hpd_w[0].shape[1] != 2:
        raise GngfError("the HPD input must be 2-D grid-corner coordinates")
    state.hpd_small = (not _streaming_ok(cfg, U, T, k, hpd_w[-1].shape[1])) and _hpd_small_ok(hpd_w, U, k)
    if state.hpd_small:
        # one launch: all layers + softmax + top-k, 8 nodes per CTA
        widths = [2] + [w.shape[0] for w in hpd_w]
        state.hpd_acts = [torch.empty((U, w.shape[0]), dtype=torch.float32, device=device) for w in hpd_w[:-1]]
        state.uprobs = torch.empty((U, T), dtype=torch.float32, device=device)
        state.utopv = torch.empty((U, k), dtype=torch.float32, device=device)
        state.utopi = torch.empty((U, k), dtype=torch.int32, device=device)
        if tables is not None and SMALL_FUSE_NODE_PASSES:
            state.nfeat = torch.empty((lat.num_level_nodes, cfg.feature_dim), dtype=torch.float32, device=device)
            call("gngf_hpd_small_fwd_enc", lat, n, _lib.int_array(widths), _lib.ptr_array(hpd_w), _lib.ptr_array(hpd_b),
                 _lib.ptr_array(state.hpd_acts), k, state.uprobs.data_ptr(), state.utopv.data_ptr(),
                 state.utopi.data_ptr(), make_tables(tables), cfg.feature_dim, cfg.mix_mode, state.nfeat.data_ptr(),
                 _stream())
            return
        call("gngf_hpd_small_fwd", lat, n, _lib.int_array(widths), _lib.ptr_array(hpd_w), _lib.ptr_array(hpd_b),
             _lib.ptr_array(state.hpd_acts), k, state.uprobs.data_ptr(), state.utopv.data_ptr(),
             state.utopi.data_ptr(), _stream())
        return
    acts = []
    ids = None
    streaming = n > 1 and _streaming_ok(cfg, U, T, k, hpd_w[-1].shape[1])
    # (node parallelism reads the node count back and sizes collectives with it: eager steps only)
    shard = NODE_SHARD if (streaming and STREAM_BWD_FUSED and not torch.cuda.is_current_stream_capturing()) else None
    if streaming and _active_nodes_wanted(U):
        ids = active_nodes(state.x, lat, shard)
        if ids.shape[0] == 0 or (shard is None and FORCE_ACTIVE_NODES is None
                                 and ids.shape[0] > ACTIVE_NODES_MAX_FRACTION * U):
            ids = None
    if shard is not None:
        # node parallelism: this rank evaluates rows r0..r1 of the node list every rank agrees on
        total = U if ids is None else int(ids.shape[0])
        r0, r1, chunk = shard.bounds(total)
        state.shard = (shard, total, r0, r1, chunk)
        state.node_ids_all = ids
        ids = torch.arange(r0, r1, dtype=torch.int32, device=device) if ids is None else ids[r0:r1]
    state.node_ids = ids
    rows = U if ids is None else ids.shape[0]
    h = torch.empty((rows, hpd_w[0].shape[0]), dtype=torch.float32, device=device)
    if rows > 0:
        call("gngf_hpd_first_layer_fwd_nodes", lat, _ptr(ids), rows, hpd_w[0].data_ptr(), hpd_b[0].data_ptr(),
             hpd_w[0].shape[0], ACT_RELU if n > 1 else ACT_NONE, h.data_ptr(), _stream())
    for i in range(1, n - 1):
        acts.append(h)
        h = linear_fwd(h, hpd_w[i], hpd_b[i], ACT_RELU) if rows > 0 else \
            torch.empty((0, hpd_w[i].shape[0]), dtype=torch.float32, device=device)
    state.hpd_acts = acts
    if n == 1:                      # single-layer HPD: h already holds the logits
        state.uprobs, state.utopv, state.utopi = softmax_topk_fwd(h, k, inplace=True)
        return
    acts.append(h)
    kd = h.shape[1]
    if _streaming_ok(cfg, U, T, k, kd):
        state.w_planes = split_f16x2(hpd_w[-1])
        state.h_planes = split_f16x2(h) if rows > 0 else None
        state.utopv, state.utopi, state.row_max, state.row_sum = hpd_stream_fwd(
            h, hpd_w[-1], hpd_b[-1], k, h_planes=state.h_planes, w_planes=state.w_planes,
            alloc_rows=None if state.shard is None else state.shard[4])
        state.uprobs = None
        if state.shard is not None:
            # every rank's selections, in list order (32 bytes per node for K = 4)
            sh, total = state.shard[0], state.shard[1]
            topv_all = sh.all_gather_rows(state.utopv, total)
            topi_all = sh.all_gather_rows(state.utopi, total)
            state.utopv_rows, state.utopi_rows = state.utopv[:rows], state.utopi[:rows]
            if state.node_ids_all is None:          # the list is the whole box
                state.utopv, state.utopi = topv_all, topi_all
            else:
                state.utopv = torch.ones((U, k), dtype=torch.float32, device=device)
                state.utopi = torch.zeros((U, k), dtype=torch.int32, device=device)
                scatter_node_rows(state.node_ids_all, topv_all, state.utopv)
                scatter_node_rows(state.node_ids_all, topi_all, state.utopi)
        elif ids is not None:
            # per-node results back into (U,K) arrays for the gather / scatter kernels.  Untouched nodes: slot 0 with
            # probability 1 -- finite under every mix mode; their multiplicities and feature adjoints are zero
            state.utopv_rows, state.utopi_rows = state.utopv, state.utopi
            state.utopv = torch.ones((U, k), dtype=torch.float32, device=device)
            state.utopi = torch.zeros((U, k), dtype=torch.int32, device=device)
            scatter_node_rows(ids, state.utopv_rows, state.utopv)
            scatter_node_rows(ids, state.utopi_rows, state.utopi)
        return
    if U * T * 4 > MAX_DENSE_LOGIT_BYTES:
        raise GngfError(f"dense logits for U={U} nodes x T={T} slots need {U * T * 4 / 2**30:.0f} GiB; the streaming "
                        "path needs should_keep_topk_only=True, topk_k <= 8 and a last hidden width <= 128")
    if U * T >= TC_MIN_ELEMENTS and kd % 8 == 0:
        logits = tc_linear_fwd(h, hpd_w[-1], hpd_b[-1], ACT_NONE)
    else:
        logits = linear_fwd(h, hpd_w[-1], hpd_b[-1], ACT_NONE)
    state.uprobs, state.utopv, state.utopi = softmax_topk_fwd(logits, k, inplace=True)


class GNGFPath(torch.autograd.Function):
    """forward + backward of GeneralNeuralGaugeFields (models.py:394-484) as one autograd node.

    inputs : x (P,2), a ForwardState shell (non-tensor), then the parameters
             [hpd_w0, hpd_b0, ..., tables 0..L-1, mlp_w0, mlp_b0, ...]
    outputs: rgb (P,C), colsum (L,N)  -- the column sums of the (virtual) `probs` tensor the reference
             returns (N = T, or K when should_keep_topk_only) -- and uvals (U,N), the per-node rows of that
             tensor (only used when a caller materialises `probs`).
    """

    @staticmethod
    def forward(ctx, x, state: ForwardState, *params):
        cfg, lat = state.cfg, state.lat
        ctx.set_materialize_grads(False)      # unused outputs (uvals, colsum) arrive as None, not as zero tensors
        dev = x.device
        L, F, K, T = lat.num_levels, cfg.feature_dim, cfg.topk_k, cfg.table_size
        P = x.shape[0]
        nh, nm = (0 if cfg.use_hash else cfg.n_hpd), cfg.n_mlp
        hpd_w, hpd_b = list(params[0:2 * nh:2]), list(params[1:2 * nh:2])
        tables = list(params[2 * nh:2 * nh + L])
        mlp_w, mlp_b = list(params[2 * nh + L::2]), list(params[2 * nh + L + 1::2])
        st = _stream()
        tab = make_tables(tables)
        enc = torch.empty((P, L * F), dtype=torch.float32, device=dev)

        if cfg.use_hash:
            call("gngf_encode_hash_fwd", x.data_ptr(), P, lat, tab, T, F, enc.data_ptr(), None, st)
            colsum = uvals = None
        else:
            hpd_forward_nodes(lat, hpd_w, hpd_b, K, dev, cfg, state, tables=tables)
            # API output idx_topk (P,L,4,K) int64: nothing in the step reads it -> side stream, joined at the end
            fork_idx = None
            if WANT_IDX_TOPK:
                state.idx_topk = torch.empty((P, L, 4, K), dtype=torch.int64, device=dev)
                fork_idx = _Fork(dev, 0)
                with fork_idx:
                    gather_rows(x, lat, state.utopi, out=state.idx_topk)
            S = lat.num_level_nodes
            nfeat, state.nfeat = state.nfeat, None       # (not kept for the backward)
            if nfeat is None:
                nfeat = torch.empty((S, F), dtype=torch.float32, device=dev)
                call("gngf_node_features_fwd", lat, tab, T, F, K, cfg.mix_mode, state.utopv.data_ptr(),
                     state.utopi.data_ptr(), nfeat.data_ptr(), st)
            # cnt (S) | cell counts (S) | err flag (1) | colsum (L*N) share one zero-initialised buffer: one memset
            N = K if cfg.topk_only else T
            # (the column sums start on a 16-byte boundary: the peer all-reduce of dp.py moves them as float4)
            c0 = (2 * S + 1 + 3) & ~3
            zbuf = torch.zeros(c0 + L * N, dtype=torch.int32, device=dev)
            state.cnt, cell_cnt = zbuf[:S], zbuf[S:2 * S]
            if state.err_flag is None:        # (a caller with fixed coordinate bounds passes its sticky flag instead)
                state.err_flag = zbuf[2 * S:2 * S + 1]
            colsum = zbuf[c0:].view(torch.float32).view(L, N)
            call("gngf_encode_fwd", x.data_ptr(), P, lat, F, nfeat.data_ptr(), enc.data_ptr(), state.cnt.data_ptr(),
                 cell_cnt.data_ptr(), state.err_flag.data_ptr(), st)
            uvals = state.utopv if cfg.topk_only else state.uprobs
            fork_col = _Fork(dev, 1)      # only the loss reads the column sums: overlaps the decoder
            with fork_col:
                # node multiplicities from the per-cell counts of the point pass (one atomic per (point, level))
                call("gngf_cell_to_node_counts", lat, cell_cnt.data_ptr(), state.cnt.data_ptr(), _stream())
                call("gngf_lattice_colsum", lat, state.cnt.data_ptr(), uvals.data_ptr(), N, colsum.data_ptr(), _stream())
                if COLSUM_REDUCE_HOOK is not None:
                    state.colsum_world = int(COLSUM_REDUCE_HOOK(colsum))

        state.mlp_fused = _mlp3_supported(mlp_w)
        if state.mlp_fused:
            # K6: one kernel, hidden activations never leave the SM (recomputed in backward)
            C = mlp_w[2].shape[0]
            rgb_out = torch.empty((P, C), dtype=torch.float32, device=dev)
            state.mlp_tc = bool(MLP_TENSOR_CORES)
            if state.mlp_tc:
                # (P,4) words: the ReLU pattern of the two hidden layers, which the backward gates with
                state.mlp_masks = torch.empty((P, 4), dtype=torch.int32, device=dev)
                call("gngf_mlp3_tc_fwd", enc.data_ptr(), P, L * F, C, int(cfg.leaky), mlp_w[0].data_ptr(),
                     mlp_b[0].data_ptr(), mlp_w[1].data_ptr(), mlp_b[1].data_ptr(), mlp_w[2].data_ptr(),
                     mlp_b[2].data_ptr(), rgb_out.data_ptr(), state.mlp_masks.data_ptr(), st)
            else:
                call("gngf_mlp3_fwd", enc.data_ptr(), P, L * F, C, int(cfg.leaky), mlp_w[0].data_ptr(),
                     mlp_b[0].data_ptr(), mlp_w[1].data_ptr(), mlp_b[1].data_ptr(), mlp_w[2].data_ptr(),
                     mlp_b[2].data_ptr(), rgb_out.data_ptr(), st)
            acts = [enc, rgb_out]
        else:
            acts = [enc]
            h = enc
            hidden_act = ACT_LEAKY_RELU if cfg.leaky else ACT_RELU
            for i in range(nm):
                h = linear_fwd(h, mlp_w[i], mlp_b[i], hidden_act if i < nm - 1 else ACT_SIGMOID)
                acts.append(h)
        if not cfg.use_hash:
            if fork_idx is not None:
                fork_idx.join()
            if DEFER_COLSUM_JOIN and fork_col.on:
                state.colsum_fork = fork_col
            else:
                fork_col.join()
        state.mlp_acts = acts
        state.x = x
        ctx.state = state
        ctx.params = params
        ctx.n_params = len(params)
        # return fresh aliases: autograd attaches grad_fn to the returned objects, and the tensors kept in
        # `state` must not keep the autograd graph (and the parameters' AccumulateGrad nodes) alive
        rgb = acts[-1].detach()
        if cfg.use_hash:
            return rgb
        return rgb, colsum.detach(), uvals.detach()

    @staticmethod
    def backward(ctx, grad_rgb, grad_colsum=None, grad_uvals=None):
        state: ForwardState = ctx.state
        cfg, lat, x = state.cfg, state.lat, state.x
        params = ctx.params
        dev = x.device
        L, F, K, T = lat.num_levels, cfg.feature_dim, cfg.topk_k, cfg.table_size
        P = x.shape[0]
        nh, nm = (0 if cfg.use_hash else cfg.n_hpd), cfg.n_mlp
        hpd_w = list(params[0:2 * nh:2])
        tables = list(params[2 * nh:2 * nh + L])
        mlp_w, mlp_b = list(params[2 * nh + L::2]), list(params[2 * nh + L + 1::2])
        st = _stream()
        S = 0 if cfg.use_hash else lat.num_level_nodes
        U = lat.num_nodes

        # one zero-initialised buffer for every accumulated gradient
        sizes = [p.numel() for p in params] + [S * F, 0 if cfg.use_hash else U * K]
        # (every view starts on a 16-byte boundary: the scatter kernels use vector reductions)
        padded = [(n + 3) & ~3 for n in sizes]
        flat = torch.zeros(sum(padded), dtype=torch.float32, device=dev)
        views, off = [], 0
        for n, npad in zip(sizes, padded):
            views.append(flat[off:off + n])
            off += npad
        grads = [v.view(p.shape) for v, p in zip(views[:len(params)], params)]
        dnf, dtv = views[-2], views[-1]
        # what the data-parallel exchange moves: the parameter gradients, not the per-node scratch behind them
        flat_params = flat[:sum(padded[:len(params)])]
        g_hpd_w, g_hpd_b = grads[0:2 * nh:2], grads[1:2 * nh:2]
        g_tables = grads[2 * nh:2 * nh + L]
        g_mlp_w, g_mlp_b = grads[2 * nh + L::2], grads[2 * nh + L + 1::2]

        # decoder MLP (models.py:468-470)
        acts = state.mlp_acts
        rgb = acts[-1]
        grad_rgb = torch.zeros_like(rgb) if grad_rgb is None else _f32c(grad_rgb)
        if state.mlp_fused:
            C = rgb.shape[1]
            denc = torch.empty((P, L * F), dtype=torch.float32, device=dev)
            if state.mlp_tc:
                call("gngf_mlp3_tc_bwd", acts[0].data_ptr(), rgb.data_ptr(), grad_rgb.data_ptr(), P, L * F, C,
                     int(cfg.leaky), mlp_w[0].data_ptr(), mlp_b[0].data_ptr(), mlp_w[1].data_ptr(), mlp_b[1].data_ptr(),
                     mlp_w[2].data_ptr(), state.mlp_masks.data_ptr(), denc.data_ptr(), g_mlp_w[0].data_ptr(),
                     g_mlp_b[0].data_ptr(), g_mlp_w[1].data_ptr(), g_mlp_b[1].data_ptr(), g_mlp_w[2].data_ptr(),
                     g_mlp_b[2].data_ptr(), st)
            else:
                work = torch.empty(_lib.load().gngf_mlp3_bwd_workspace_floats(L * F, C), dtype=torch.float32,
                                   device=dev)
                call("gngf_mlp3_bwd", acts[0].data_ptr(), grad_rgb.data_ptr(), P, L * F, C, int(cfg.leaky),
                     mlp_w[0].data_ptr(), mlp_b[0].data_ptr(), mlp_w[1].data_ptr(), mlp_b[1].data_ptr(),
                     mlp_w[2].data_ptr(), mlp_b[2].data_ptr(), denc.data_ptr(), g_mlp_w[0].data_ptr(),
                     g_mlp_b[0].data_ptr(), g_mlp_w[1].data_ptr(), g_mlp_b[1].data_ptr(), g_mlp_w[2].data_ptr(),
                     g_mlp_b[2].data_ptr(), work.data_ptr(), st)
        else:
            dz = torch.empty_like(rgb)
            call("gngf_sigmoid_bwd", grad_rgb.data_ptr(), rgb.data_ptr(), rgb.numel(), dz.data_ptr(), st)
            hidden_act = ACT_LEAKY_RELU if cfg.leaky else ACT_RELU
            for i in range(nm - 1, -1, -1):
                dz = linear_bwd(dz, acts[i], mlp_w[i], hidden_act if i > 0 else ACT_NONE, True, g_mlp_w[i], g_mlp_b[i])
            denc = dz

        gtab = make_tables(g_tables)
        if cfg.use_hash:
            call("gngf_encode_hash_bwd", x.data_ptr(), P, lat, gtab, T, F, denc.data_ptr(), st)
            if GRAD_REDUCE_HOOK is not None:
                GRAD_REDUCE_HOOK(flat_params)
            return (None, None, *grads)

        call("gngf_encode_bwd", x.data_ptr(), P, lat, F, denc.data_ptr(), dnf.data_ptr(), st)
        if state.colsum_fork is not None:      # deferred join: the column sums' consumers (and grad_colsum) live on the
            state.colsum_fork.join()           # side stream up to here
            state.colsum_fork = None
        need_hpd = cfg.hpd_trainable
        fuse_nodes = need_hpd and state.hpd_small and SMALL_FUSE_NODE_PASSES and not cfg.drop_topk_adjoint
        if not fuse_nodes:
            call("gngf_node_features_bwd", lat, make_tables(tables), gtab, T, F, K, cfg.mix_mode, state.utopv.data_ptr(),
                 state.utopi.data_ptr(), dnf.data_ptr(), dtv.data_ptr() if need_hpd else None, st)
        if not need_hpd:
            for i in range(2 * nh):
                grads[i] = None
            if GRAD_REDUCE_HOOK is not None:
                GRAD_REDUCE_HOOK(flat_params)
            return (None, None, *grads)

        gcol = gcol_k = gdense = None
        dtv_extra = False            # dtv holds an adjoint of utopv that did not come from the node pass
        if cfg.drop_topk_adjoint:
            # reference semantics of should_inplace_scatter = None: every adjoint that would pass through the top-k values
            # (the mix weights and, in top-k-only mode, the returned probabilities) is dropped
            dtv.zero_()
            if cfg.topk_only:
                grad_colsum = grad_uvals = None
        if grad_colsum is not None:
            grad_colsum = _f32c(grad_colsum)
            if state.colsum_world > 1:
                grad_colsum = grad_colsum * float(state.colsum_world)
            if cfg.topk_only:
                gcol_k = grad_colsum
            else:
                gcol = grad_colsum
        if grad_uvals is not None:
            grad_uvals = _f32c(grad_uvals)
            if cfg.topk_only:
                dtv.add_(grad_uvals.reshape(-1))
                dtv_extra = True
            else:
                gdense = grad_uvals
        if state.hpd_small:
            # one launch: dlogits, dX chain with ReLU masks, bias and first-layer gradients; then the weight
            # gradients dW_i += g_i^T h_{i-1} (reductions over all nodes) as split-K layer kernels
            widths = [2] + [w.shape[0] for w in hpd_w]
            gacts = [torch.empty((U, w.shape[0]), dtype=torch.float32, device=dev) for w in hpd_w]
            if fuse_nodes:
                # ... and the encoding's node pass: table gradients, adjoint of the selected probabilities in registers
                call("gngf_hpd_small_bwd_enc", lat, nh, _lib.int_array(widths), _lib.ptr_array(hpd_w),
                     _lib.ptr_array(state.hpd_acts), _lib.ptr_array(gacts), _lib.ptr_array(g_hpd_b),
                     g_hpd_w[0].data_ptr(), K, state.uprobs.data_ptr(), state.utopi.data_ptr(),
                     dtv.data_ptr() if dtv_extra else None, state.cnt.data_ptr(), _ptr(gcol), _ptr(gcol_k), _ptr(gdense),
                     make_tables(tables), gtab, F, cfg.mix_mode, dnf.data_ptr(), st)
            else:
                call("gngf_hpd_small_bwd", lat, nh, _lib.int_array(widths), _lib.ptr_array(hpd_w),
                     _lib.ptr_array(state.hpd_acts), _lib.ptr_array(gacts), _lib.ptr_array(g_hpd_b), g_hpd_w[0].data_ptr(),
                     K, state.uprobs.data_ptr(), state.utopi.data_ptr(), dtv.data_ptr(), state.cnt.data_ptr(), _ptr(gcol),
                     _ptr(gcol_k), _ptr(gdense), st)
            forks = []
            for i in range(1, nh):      # independent products: one stream each
                f = _Fork(dev, i - 1)
                with f:
                    call("gngf_linear_bwd", gacts[i].data_ptr(), state.hpd_acts[i - 1].data_ptr(), hpd_w[i].data_ptr(), U,
                         hpd_w[i].shape[0], hpd_w[i].shape[1], ACT_NONE, None, g_hpd_w[i].data_ptr(), None, _stream())
                forks.append(f)
            for f in forks:
                f.join()
            if GRAD_REDUCE_HOOK is not None:
                GRAD_REDUCE_HOOK(flat_params)
            return (None, None, *grads)
        if nh == 1:
            dlogits = torch.empty((U, T), dtype=torch.float32, device=dev)
            call("gngf_hpd_dlogits", lat, state.uprobs.data_ptr(), T, K, state.utopi.data_ptr(), dtv.data_ptr(),
                 state.cnt.data_ptr(), _ptr(gcol), _ptr(gcol_k), _ptr(gdense), None, None, 0, U, dlogits.data_ptr(), st)
            dz = dlogits
        elif state.uprobs is not None:
            dlogits = torch.empty((U, T), dtype=torch.float32, device=dev)
            call("gngf_hpd_dlogits", lat, state.uprobs.data_ptr(), T, K, state.utopi.data_ptr(), dtv.data_ptr(),
                 state.cnt.data_ptr(), _ptr(gcol), _ptr(gcol_k), _ptr(gdense), None, None, 0, U, dlogits.data_ptr(), st)
            dz = linear_bwd(dlogits, state.hpd_acts[nh - 2], hpd_w[nh - 1], ACT_RELU, True, g_hpd_w[nh - 1],
                            g_hpd_b[nh - 1])
        elif STREAM_BWD_FUSED:
            # streaming path, fused: logits are recomputed tile by tile in TMEM, turned into dlogits by the epilogue
            # warps and contracted again (dlogits W3, dlogits^T h) without leaving the SM (k2_hpd_tc_bwd.cu)
            h_last = state.hpd_acts[nh - 2]
            kd = h_last.shape[1]
            ids = state.node_ids
            if state.shard is not None:
                # node parallelism: this rank's share of the adjoints of ALL listed nodes (column-sum adjoint folded in),
                # summed over ranks and handed to the rows' owners; the owner's backward takes them row-indexed
                sh, total, r0, r1, chunk = state.shard
                adj = torch.zeros((sh.world * chunk, K), dtype=torch.float32, device=dev)
                call("gngf_gather_node_adjoints", lat, _ptr(state.node_ids_all), total, K, dtv.data_ptr(),
                     _ptr(state.cnt), _ptr(gcol_k), adj.data_ptr(), st)
                mine = sh.reduce_scatter_rows(adj)
                del adj
                if r1 > r0:
                    dz = hpd_stream_bwd(lat, h_last, hpd_w[nh - 1], params[2 * (nh - 1) + 1], state.h_planes,
                                        state.w_planes, state.utopv_rows, state.utopi_rows, mine, None, None,
                                        state.row_max, state.row_sum, g_hpd_w[nh - 1], g_hpd_b[nh - 1], node_ids=None)
                else:
                    dz = torch.empty((0, kd), dtype=torch.float32, device=dev)
            else:
                dz = hpd_stream_bwd(lat, h_last, hpd_w[nh - 1], params[2 * (nh - 1) + 1], state.h_planes, state.w_planes,
                                    state.utopv if ids is None else state.utopv_rows,
                                    state.utopi if ids is None else state.utopi_rows, dtv, state.cnt, gcol_k,
                                    state.row_max, state.row_sum, g_hpd_w[nh - 1], g_hpd_b[nh - 1], node_ids=ids)
        else:
            # streaming path: recompute the logits chunk by chunk on the tensor cores, turn them into dlogits in
            # place from the saved softmax statistics, and feed the output layer's backward
            h_last = state.hpd_acts[nh - 2]
            kd = h_last.shape[1]
            rows = int(max(128, min(U, BWD_CHUNK_BYTES // (T * 4)))) // 8 * 8
            dz = torch.zeros((U, kd), dtype=torch.float32, device=dev)
            wt_planes = split_bf16x3(hpd_w[nh - 1].t().contiguous())          # (3, kd, T): B operand of dX
            wb_planes = split_bf16x3(hpd_w[nh - 1])                           # (3, T, kd): B operand of the logits
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            for r0 in range(0, U, rows):
                n = min(rows, U - r0)
                hc = h_last[r0:r0 + n]
                buf = tc_linear_fwd(hc, hpd_w[nh - 1], params[2 * (nh - 1) + 1], ACT_NONE, w_planes=wb_planes)
                call("gngf_hpd_dlogits", lat, buf.data_ptr(), T, K, state.utopi.data_ptr(), dtv.data_ptr(),
                     state.cnt.data_ptr(), None, _ptr(gcol_k), None, state.row_max.data_ptr(),
                     state.row_sum.data_ptr(), r0, n, buf.data_ptr(), st)
                # db += colsum(dlogits); dX = dlogits W3 (masked); dW3 += dlogits^T h  -- all on tcgen05
                call("gngf_linear_bwd", buf.data_ptr(), hc.data_ptr(), hpd_w[nh - 1].data_ptr(), n, T, kd, ACT_NONE,
                     None, None, g_hpd_b[nh - 1].data_ptr(), st)
                dl_planes = split_bf16x3(buf)                                   # (3, n, T)
                dxc = dz[r0:r0 + n]
                out_tiles = ((n + 127) // 128) * ((kd + 127) // 128)
                tc_gemm_planes(dl_planes, wt_planes, dxc, k_splits=max(1, (2 * sms) // out_tiles))
                dxc.mul_(hc > 0)
                del dl_planes
                dlt_planes = split_bf16x3_t(buf)                                # (3, T, n8)
                ht_planes = split_bf16x3_t(hc)                                  # (3, kd, n8)
                tc_gemm_planes(dlt_planes, ht_planes, g_hpd_w[nh - 1], accumulate=True)
                del buf, dlt_planes, ht_planes
        if dz.shape[0] > 0:          # (a node-parallel rank can own no rows of a tiny list)
            for i in range(nh - 2, 0, -1):
                dz = linear_bwd(dz, state.hpd_acts[i - 1], hpd_w[i], ACT_RELU, True, g_hpd_w[i], g_hpd_b[i])
            call("gngf_hpd_first_layer_bwd_nodes", lat, _ptr(state.node_ids), dz.shape[0], dz.data_ptr(),
                 hpd_w[0].shape[0], g_hpd_w[0].data_ptr(), g_hpd_b[0].data_ptr(), st)
        if GRAD_REDUCE_HOOK is not None:
            GRAD_REDUCE_HOOK(flat_params)      # every parameter gradient is a view of it: one collective for all
        return (None, None, *grads)


class GatherRows(torch.autograd.Function):
    """Materialises (P,L,4,N) rows of a per-node array (the slow path behind LazyProbs)."""

    @staticmethod
    def forward(ctx, uvals, x, lat):
        ctx.x, ctx.lat, ctx.shape = x, lat, uvals.shape
        return gather_rows(x, lat, uvals)

    @staticmethod
    def backward(ctx, grad_out):
        grad_out = _f32c(grad_out)
        dvals = torch.zeros(ctx.shape, dtype=torch.float32, device=grad_out.device)
        call("gngf_lattice_scatter_rows", ctx.x.data_ptr(), ctx.x.shape[0], ctx.lat, grad_out.data_ptr(), ctx.shape[1],
             dvals.data_ptr(), _stream())
        return dvals, None, None
