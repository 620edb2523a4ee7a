"""Drop-in replacement for the reference's ``models.py``: same four classes, constructor arguments, forward
signatures, return values and state-dict keys (reference models.py:5-656), with the training hot path
executed by the sm_100a CUDA library behind include/gngf.h instead of ~760 ATen ops per step.

``main.py`` / ``functions.py`` of the reference run unchanged when this module is importable as ``models``
ahead of the reference's own (the repo-root ``models.py`` shim does that): ``main.py:3`` star-imports the
classes and hands ``GeneralNeuralGaugeFields`` to ``grid_search_loop`` (functions.py:540-556).

Global flags.  The reference reads ``should_use_hash_function``, ``should_softmax_topk_features``,
``should_inplace_scatter``, ``should_leaky_relu`` and ``should_batchnorm_data`` as module globals star-imported
from ``params`` (models.py:1-2).  Here they are looked up at call time in the ``params`` module when one is
importable (i.e. under the reference's driver) and otherwise in :data:`DEFAULT_FLAGS` (= params.py:1-23).
"""
from __future__ import annotations

import sys
from collections import Counter
from types import SimpleNamespace
from typing import Tuple

import numpy as np
import torch
import torch.nn as nn

from . import ops
from ._lib import GngfError
from .lattice import build_lattice, level_resolutions
from .lazy_probs import LazyProbs

__all__ = ["DifferentiableTopk", "HashProbDistribution", "MultiResHashEncoding", "GeneralNeuralGaugeFields"]

# params.py:1-23
DEFAULT_FLAGS = SimpleNamespace(
    should_batchnorm_data=False,
    should_inplace_scatter=True,
    should_softmax_topk_features=True,
    should_leaky_relu=False,
    should_use_hash_function=False,
)
_FLAG_NAMES = tuple(vars(DEFAULT_FLAGS))


def current_flags() -> SimpleNamespace:
    src = sys.modules.get("params")
    if src is None:
        return DEFAULT_FLAGS
    return SimpleNamespace(**{n: getattr(src, n, getattr(DEFAULT_FLAGS, n)) for n in _FLAG_NAMES})


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise GngfError("no CUDA device: collision_handling_in_instantngp_b200 has no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


class DifferentiableTopk(torch.autograd.Function):
    """reference models.py:5-42.  forward: (values, int64 indices) of the k largest entries, sorted, ties
    towards the lower index; backward: scatter of grad_values into zeros."""

    @staticmethod
    def forward(ctx, input: torch.Tensor, k: int, dim: int):
        ctx.dim = dim
        moved = input.movedim(dim, -1)
        ctx.in_shape = moved.shape
        values, indices = ops.topk_fwd(moved, k)
        ctx.save_for_backward(indices)
        ctx.mark_non_differentiable(indices)
        return values.movedim(-1, dim), indices.movedim(-1, dim)

    @staticmethod
    def backward(ctx, grad_values: torch.Tensor, grad_indices: torch.Tensor):
        (indices,) = ctx.saved_tensors
        # should_inplace_scatter (models.py:30-35): True / False select between equivalent scatter variants; None calls
        # the out-of-place scatter and DISCARDS its result (models.py:30-31), so grad_input stays all zeros
        if current_flags().should_inplace_scatter is None:
            return torch.zeros(ctx.in_shape, dtype=grad_values.dtype, device=grad_values.device).movedim(-1, ctx.dim), None, None
        grad_in = ops.topk_bwd(grad_values.movedim(ctx.dim, -1), indices, ctx.in_shape[-1])
        return grad_in.movedim(-1, ctx.dim), None, None


class _Linear(torch.autograd.Function):
    """y = act(x w^T + b) through gngf_linear_fwd / gngf_linear_bwd (used by the stand-alone module forwards)."""

    @staticmethod
    def forward(ctx, x, w, b, act):
        shape = x.shape
        x2 = ops._f32c(x).reshape(-1, shape[-1])
        y = ops.linear_fwd(x2, w, b, act)
        ctx.save_for_backward(x2, w, y)
        ctx.act, ctx.shape = act, shape
        return y.reshape(*shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, gy):
        x2, w, y = ctx.saved_tensors
        gy = ops._f32c(gy).reshape(-1, w.shape[0])
        if ctx.act == ops.ACT_SIGMOID:
            gy = gy * y * (1 - y)
        elif ctx.act == ops.ACT_RELU:
            gy = gy * (y > 0)
        elif ctx.act == ops.ACT_LEAKY_RELU:
            gy = torch.where(y > 0, gy, gy * 0.01)
        gy = gy.contiguous()
        dw, db = torch.zeros_like(w), torch.zeros(w.shape[0], dtype=torch.float32, device=w.device)
        dx = ops.linear_bwd(gy, x2, w, ops.ACT_NONE, True, dw, db)
        return dx.reshape(ctx.shape), dw, db, None


class _SoftmaxNanToNum(torch.autograd.Function):
    """nan_to_num(softmax(z, -1)) (models.py:85,111)."""

    @staticmethod
    def forward(ctx, z):
        shape = z.shape
        probs, _, _ = ops.softmax_topk_fwd(ops._f32c(z).reshape(-1, shape[-1]), 1)
        probs = probs.reshape(shape)
        ctx.save_for_backward(probs)
        return probs

    @staticmethod
    def backward(ctx, g):
        (p,) = ctx.saved_tensors
        return p * (g - (g * p).sum(-1, keepdim=True))


class HashProbDistribution(nn.Module):
    """The GNGF h(x): MLP [in, *hidden, out] with ReLU between layers and a softmax on top
    (reference models.py:45-123).  Inside GeneralNeuralGaugeFields it is evaluated on lattice nodes by the fused
    path; this forward keeps the stand-alone API: (probs, topk_probs, topk_indices)."""

    def __init__(self, hidden_layers_widths: list, in_features: int = 2, out_features: int = 2 ** 14, k: int = 1,
                 topk_dim: int = -1, should_log: bool = False):
        super().__init__()
        self.out_features = out_features
        self._k = k
        self._topk_dim = topk_dim
        self._should_log = should_log
        widths = [in_features, *hidden_layers_widths, out_features]
        dev = _device()
        # same module tree (and RNG consumption order) as models.py:82-88 -> identical state-dict keys
        self.module_list = nn.ModuleList([
            nn.Sequential(
                nn.Linear(widths[i], widths[i + 1], device=dev),
                nn.ReLU() if i < len(widths) - 2 else nn.Softmax(dim=-1))
            for i in range(len(widths) - 1)
        ])

    def weights(self):
        return [m[0].weight for m in self.module_list], [m[0].bias for m in self.module_list]

    def forward(self, x: torch.Tensor) -> Tuple:
        ws, bs = self.weights()
        h = x
        for i, (w, b) in enumerate(zip(ws, bs)):
            h = _Linear.apply(h, w, b, ops.ACT_RELU if i < len(ws) - 1 else ops.ACT_NONE)
        probs = _SoftmaxNanToNum.apply(h).squeeze(-1)
        topk_probs, topk_indices = DifferentiableTopk.apply(probs, self._k, self._topk_dim)
        if self._k != 1:
            topk_probs = topk_probs.squeeze(-1)
            topk_indices = topk_indices.squeeze(-1)
        return probs, topk_probs, topk_indices


class MultiResHashEncoding(nn.Module):
    """L feature tables nn.Embedding(T, F), U(-1e-4, 1e-4) (reference models.py:126-236)."""

    def __init__(self, hash_table_size: int, num_levels: int, feature_dim: int = 2, topk_k: int = 4,
                 should_log: bool = False) -> None:
        super().__init__()
        self._hash_table_size = hash_table_size
        self._num_levels = num_levels
        self._feature_dim = feature_dim
        self._topk_k = topk_k
        self._should_log = should_log
        dev = _device()
        self._hash_tables = nn.ModuleList(
            [nn.Embedding(hash_table_size, feature_dim, device=dev) for _ in range(num_levels)])
        self._apply_init(nn.init.uniform_, -10.0 ** (-4), 10.0 ** (-4))

    def tables(self):
        return [t.weight for t in self._hash_tables]

    def forward(self, hashed_indices: torch.Tensor, hashed_probs_topk: torch.Tensor, should_calc_counts: bool = False):
        """Stand-alone API (models.py:173-229): per-row indices (P,L,4,K) [or (P,L,4) in hash mode] and top-k
        probabilities -> (P,F,L,4).  Compatibility path (torch CUDA ops); the training path never calls it."""
        flags = current_flags()
        tabs = self.tables()
        if flags.should_use_hash_function:
            looked = torch.stack([tabs[j][hashed_indices[:, j].long()] for j in range(self._num_levels)], dim=1)
            return looked.permute(0, 3, 1, 2)                                        # (P,F,L,4)
        g = torch.stack([tabs[j][hashed_indices[:, j].long()] for j in range(self._num_levels)], dim=1)  # (P,L,4,K,F)
        pk = hashed_probs_topk.unsqueeze(-1)
        mode = flags.should_softmax_topk_features
        if mode is None:
            feat = (g * pk).sum(3)
        elif mode:
            feat = (g * torch.softmax(hashed_probs_topk, dim=-1).unsqueeze(-1)).sum(3)
        else:
            feat = (g * pk).sum(3) / pk.sum(3)
        return feat.permute(0, 3, 1, 2)

    def _apply_init(self, init_func, *args):
        for i in range(self._num_levels):
            init_func(self._hash_tables[i].weight, *args)


class GeneralNeuralGaugeFields(nn.Module):
    """reference models.py:239-656 -- constructor arguments, attributes, forward returns and helper methods
    are the reference's; forward/backward run through :class:`ops.GNGFPath`."""

    def __init__(self, input_dim: list, hash_table_size: int, num_levels: int, n_min: int, n_max: int,
                 MLP_hidden_layers_widths: list, HPD_hidden_layers_widths: list, HPD_out_features: int = 1,
                 feature_dim: int = 2, topk_k: int = 4, should_keep_topk_only: bool = False, should_bw: bool = False,
                 should_log: bool = False, HPD_weights_path: str = None, encoding_weights_path: str = None):
        super().__init__()
        flags = current_flags()
        if input_dim != 2:
            raise GngfError("the CUDA path implements the reference's 2-D image task (input_dim == 2)")
        self._hash_table_size = hash_table_size
        self._num_levels = num_levels
        self._n_min = n_min
        self._n_max = n_max
        self._feature_dim = feature_dim
        self._input_dim = input_dim
        self._topk_k = topk_k
        self._should_log = should_log
        self._should_keep_topk_only = should_keep_topk_only
        dev = _device()

        # models.py:305-317
        b = np.exp((np.log(n_max) - np.log(n_min)) / (num_levels - 1))
        if b > 2 or b <= 1:
            print(f"The between level scale is recommended to be <= 2 and needs to be > 1 but was {b:.4f}.")
        self._n_ls_host = level_resolutions(n_min, n_max, num_levels)
        self._n_ls = torch.from_numpy(self._n_ls_host.astype(np.float64)).reshape(1, 1, -1, 1).to(dev).int()
        cube = np.stack([[0, 1, 0, 1], [0, 0, 1, 1]])                                 # models.py:322-331
        self._voxels_helper_hypercube = torch.from_numpy(cube).unsqueeze(0).unsqueeze(2).to(dev).int()

        # layers, in the reference's construction order (RNG parity of the initial weights)
        self._batch_norm = nn.BatchNorm1d(input_dim, device=dev)                      # models.py:340
        self._use_hash = bool(flags.should_use_hash_function)
        if self._use_hash:
            self._prime_numbers = nn.Parameter(torch.from_numpy(np.array([1, 2654435761, 805459861])).to(dev), False)
        else:
            if HPD_out_features != hash_table_size:
                raise GngfError("HPD_out_features must equal hash_table_size (the HPD picks table slots)")
            self.HPD = HashProbDistribution(hidden_layers_widths=HPD_hidden_layers_widths, in_features=input_dim,
                                            out_features=HPD_out_features, k=topk_k, topk_dim=-1)
            if HPD_weights_path is not None:                                          # models.py:364-371
                self.HPD.load_state_dict(torch.load(HPD_weights_path))
                print("Loaded")
                for name, param in self.HPD.named_parameters():
                    param.requires_grad = False
                    print(name, param.requires_grad)
        self.encoding = MultiResHashEncoding(hash_table_size=hash_table_size, num_levels=num_levels,
                                             feature_dim=feature_dim, topk_k=topk_k, should_log=should_log)
        self._MLP_hidden_layers_widths = [num_levels * feature_dim, *MLP_hidden_layers_widths, (3 if not should_bw else 1)]
        widths = self._MLP_hidden_layers_widths
        self.mlp = nn.ModuleList([
            nn.Sequential(
                nn.Linear(widths[i], widths[i + 1], device=dev),
                (nn.LeakyReLU() if flags.should_leaky_relu else nn.ReLU()) if i < len(widths) - 2 else nn.Sigmoid())
            for i in range(len(widths) - 1)
        ])
        self._leaky = bool(flags.should_leaky_relu)

        # coordinate bounds of the lattice: None = derive from every batch (one tiny device->host read);
        # set_coord_bounds() fixes them (CUDA-graph capture, benchmarks)
        self._coord_bounds = None
        self._lattice_cache = {}
        self.last_state = None
        # sticky device flag: a coordinate fell outside the bounds promised to set_coord_bounds() (the kernels clamp it
        # to the lattice box); read by check_errors() -- every ERR_CHECK_EVERY-th eager forward, and by GraphedTrainer
        self._err_flag_sticky = torch.zeros(1, dtype=torch.int32, device=dev)
        self._forwards = 0

    # ------------------------------------------------------------------------------------------------
    def set_coord_bounds(self, lo=(0.0, 0.0), hi=(1.0, 1.0)) -> None:
        """Promise that every coordinate passed to forward lies in [lo, hi] (per dimension)."""
        self._coord_bounds = None if lo is None else (tuple(float(v) for v in lo), tuple(float(v) for v in hi))

    ERR_CHECK_EVERY = 64

    def check_errors(self) -> None:
        """Raises GngfError if any forward since the last call saw coordinates outside the fixed bounds (one 4-byte
        device->host read; the flag is then cleared)."""
        if int(self._err_flag_sticky.item()):
            self._err_flag_sticky.zero_()
            raise GngfError(f"coordinates outside the bounds given to set_coord_bounds{self._coord_bounds} were clamped to "
                            "the lattice box: the affected forwards used the wrong lattice nodes")

    def _lattice_for(self, x: torch.Tensor):
        bounds = self._coord_bounds
        if bounds is None:
            lo, hi = torch.aminmax(x, dim=0)
            both = torch.stack([lo, hi]).tolist()
            bounds = (tuple(both[0]), tuple(both[1]))
        lat = self._lattice_cache.get(bounds)
        if lat is None:
            if len(self._lattice_cache) > 64:
                self._lattice_cache.clear()
            lat = self._lattice_cache[bounds] = build_lattice(self._n_ls_host, *bounds)
        return lat

    def _path_config(self, flags) -> ops.PathConfig:
        use_hash = self._use_hash
        return ops.PathConfig(
            table_size=self._hash_table_size, feature_dim=self._feature_dim, topk_k=self._topk_k,
            n_hpd=0 if use_hash else len(self.HPD.module_list), n_mlp=len(self.mlp),
            topk_only=self._should_keep_topk_only, mix_mode=ops.mix_mode_of(flags.should_softmax_topk_features),
            leaky=self._leaky, use_hash=use_hash, drop_topk_adjoint=flags.should_inplace_scatter is None,
            hpd_trainable=(not use_hash) and any(p.requires_grad for p in self.HPD.parameters()))

    def _parameters_flat(self):
        ps = []
        if not self._use_hash:
            for m in self.HPD.module_list:
                ps += [m[0].weight, m[0].bias]
        ps += self.encoding.tables()
        for m in self.mlp:
            ps += [m[0].weight, m[0].bias]
        return ps

    # ------------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, batch_percentage: float, should_calc_counts: bool = False):
        flags = current_flags()
        if flags.should_batchnorm_data:                                               # models.py:396-397
            x = self._batch_norm(x)
        ops._require_cuda(x, "x")
        x = ops._f32c(x.detach())
        if x.dim() != 2 or x.shape[1] != 2:
            raise GngfError(f"x must be (P, 2), got {tuple(x.shape)}")
        lat = self._lattice_for(x)
        state = ops.ForwardState(x=x, lat=lat, cfg=self._path_config(flags))
        if self._coord_bounds is not None:
            # fixed bounds: out-of-box coordinates raise the sticky flag (per-forward flag otherwise: bounds derived from
            # the batch cannot be violated)
            state.err_flag = self._err_flag_sticky
            self._forwards += 1
            if self._forwards % self.ERR_CHECK_EVERY == 0 and not torch.cuda.is_current_stream_capturing():
                self.check_errors()
        params = self._parameters_flat()
        self.last_state = state
        if self._use_hash:
            rgb = ops.GNGFPath.apply(x, state, *params)
            hashed = ops.fast_hash_fwd(x, lat, self._hash_table_size)
            counts = self._calc_counts_per_level(hashed, ops.corners_fwd(x, lat)[1]) if should_calc_counts else []
            return rgb, None, hashed, counts
        rgb, colsum, uvals = ops.GNGFPath.apply(x, state, *params)
        idx_topk = state.idx_topk                                                     # (P,L,4,K) int64
        counts = []
        if should_calc_counts:                                                        # models.py:431-439
            counts = self._calc_counts_per_level(idx_topk[..., 0], ops.corners_fwd(x, lat)[1])
        probs = LazyProbs(x, lat, uvals, colsum)
        return rgb, probs, idx_topk, counts

    # ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _scale_to_grid(self, x: torch.Tensor):
        """models.py:486-502 (materialised; the fused kernels recompute the corners in registers)."""
        ops._require_cuda(x, "x")
        x = ops._f32c(x)
        return ops.corners_fwd(x, build_lattice(self._n_ls_host))

    @torch.no_grad()
    def _fast_hash(self, grid: torch.Tensor) -> torch.Tensor:
        """models.py:504-528 on an integer corner tensor (P, 2, L, 4): the documented int32-wrap semantics."""
        gx = grid[:, 0].to(torch.int64) & 0xFFFFFFFF
        gy = ((grid[:, 1].to(torch.int64) & 0xFFFFFFFF) * 2654435761) & 0xFFFFFFFF
        h = gx ^ gy
        h = torch.where(h >= 2 ** 31, h - 2 ** 32, h)                                 # sign-extend the int32 product
        return torch.remainder(h, self._hash_table_size)

    @torch.no_grad()
    def _calc_counts_per_level(self, hash: torch.Tensor, grid: torch.Tensor):
        """Diagnostic histogram (models.py:530-566): per level {slot: number of distinct grid cells counted for it},
        with the reference's own indexing (first point of each distinct cell -> element of the flattened "(p v)" slot
        vector).  The reference de-duplicates on the host (np.unique(axis=0) per level); here two GPU passes
        (k8_collisions.cu) and one (L,T) histogram read-back.  Inputs that are not grid corners of this model's levels
        (or not CUDA tensors) take the reference's host route."""
        L = self._num_levels
        if grid.is_cuda and hash.is_cuda and grid.dim() == 4 and tuple(grid.shape[1:]) == (2, L, 4):
            # corners of x in [0,1] (main.py:48-49); wider boxes as the corners demand
            try:
                lat = self._counts_lattice(grid)
            except ValueError:                       # corners too far apart for a lattice box
                lat = None
            hist = None
            if lat is not None and lat.num_level_nodes <= 2 ** 30:
                hist, outliers = ops.counts_per_level(grid, lat, hash, self._hash_table_size)
            if hist is not None and int(outliers.item()) == 0:
                h = hist.cpu().numpy()
                out = []
                for level in range(L):
                    nz = np.nonzero(h[level])[0]
                    out.append({int(k): int(h[level][k]) for k in nz})
                return out
        P = grid.shape[0]
        rearranged = grid.permute(2, 0, 3, 1).reshape(L, P, -1).cpu().numpy()                  # "p xy l v -> l p (v xy)"
        vertices = hash.permute(1, 0, 2).reshape(L, -1).cpu().numpy()                          # "p l v -> l (p v)"
        out = []
        for level in range(L):
            _, first = np.unique(rearranged[level], axis=0, return_index=True)
            out.append(dict(Counter(vertices[level][first].tolist())))
        return out

    def _counts_lattice(self, grid: torch.Tensor):
        """Lattice whose level boxes hold every corner of `grid`: the [0,1] box of the image task, or -- when the
        coordinates were batch-normalised (params.should_batchnorm_data) -- the box of the corners themselves."""
        if self._coord_bounds is not None:
            return self._lattice_cache.get(self._coord_bounds) or build_lattice(self._n_ls_host, *self._coord_bounds)
        lo = grid[:, :, :, 0].amin(dim=0)                                                       # (2, L) floor corners
        hi = grid[:, :, :, 0].amax(dim=0)
        n = torch.from_numpy(self._n_ls_host.astype(np.float32)).to(grid.device)
        # coordinates that reproduce these corner ranges on every level: (corner + 0.5) / n_l lies inside the cell
        lo_c = ((lo + 0.5) / n).min(dim=1).values.tolist()
        hi_c = ((hi + 0.5) / n).max(dim=1).values.tolist()
        return build_lattice(self._n_ls_host, (min(lo_c[0], 0.0), min(lo_c[1], 0.0)), (max(hi_c[0], 1.0), max(hi_c[1], 1.0)))

    @torch.no_grad()
    def calc_hash_collisions(self, indices: torch.Tensor):
        """models.py:568-619: per level, (n_l+1)^2 minus the number of distinct slots in use (mean over the top-k
        columns, clamped at 0; raw integers in hash-function mode).  One bitmap pass on the GPU instead of one
        torch.unique per (column, level); values that are not integers in [0, T) -- train_step's buffer is
        torch.empty -- fall back to exact counting with torch.unique."""
        ops._require_cuda(indices, "indices")
        dev = indices.device
        L = self._num_levels
        nodes = torch.tensor([(int(n) + 1) ** 2 for n in self._n_ls_host], device=dev)
        idx4 = indices.unsqueeze(-1) if self._use_hash else indices
        uniq, outliers = ops.count_distinct(idx4, self._hash_table_size)
        if int(outliers.item()) != 0:
            cols = []
            for k in range(idx4.shape[-1]):
                per_level = idx4[..., k].permute(1, 0, 2).reshape(L, -1)
                cols.append([torch.unique(per_level[i]).shape[0] for i in range(L)])
            uniq = torch.tensor(cols, device=dev)
        if self._use_hash:
            collisions = (nodes - uniq[0]).cpu()                                      # reference: a CPU int64 tensor
        else:
            collisions = (nodes.unsqueeze(0) - uniq).float().mean(dim=0)
            collisions[collisions < 0] = 0
        minp = nodes - self._hash_table_size
        minp[minp < 0] = 0
        return collisions, minp

    def _bilinear_interpolate(self, scaled_coords, grid_coords, features):
        """models.py:621-655 on materialised tensors (compatibility API; torch CUDA ops)."""
        a, d = grid_coords[:, :, :, 0], grid_coords[:, :, :, 3]
        s = scaled_coords[:, :, :, 0]
        coeffs = torch.stack([(d[:, 0] - s[:, 0]) * (d[:, 1] - s[:, 1]), (s[:, 0] - a[:, 0]) * (d[:, 1] - s[:, 1]),
                              (d[:, 0] - s[:, 0]) * (s[:, 1] - a[:, 1]), (s[:, 0] - a[:, 0]) * (s[:, 1] - a[:, 1])], dim=-1)
        summed = (features * coeffs.unsqueeze(1)).sum(-1)                             # (P,F,L)
        return summed.permute(0, 2, 1).reshape(summed.shape[0], -1)
