"""The `probs` return value of GeneralNeuralGaugeFields.forward without its (P, L, 4, N) materialisation.

The reference returns ``hashed_probs.clone()`` -- one N-vector per (point, level, corner) row, 940 MB per batch
at the published configuration (models.py:478) -- and the only consumers are ``probs.shape`` and, in ``Loss``,
``prob[:, l, :].sum(0).sum(0) / div`` (functions.py:230, utils.py:113-116,138).  Every row equals the row of its
lattice node, so the tensor is fully described by the per-node rows ``uvals`` (U, N) and the point -> node map,
and the column sums the loss wants come straight from the node multiplicities (``gngf_lattice_colsum``).

``LazyProbs`` answers exactly those uses symbolically, with autograd attached to the column sums, and turns
into the real tensor (``materialize()``, a differentiable gather) for anything else -- indexing patterns it
does not know, attribute access, or any ``torch.*`` function called on it.
"""
from __future__ import annotations

import torch

from . import ops


class _LevelSlice:
    """prob[:, l, :]  -- virtual (P, 4, N)."""

    def __init__(self, parent: "LazyProbs", level: int):
        self._parent, self._level = parent, level
        P, _, V, N = parent.shape
        self.shape = torch.Size((P, V, N))

    def sum(self, dim=None, *args, **kwargs):
        if dim == 0 and not args and not kwargs:
            return _LevelSliceSum0(self._parent, self._level)
        return self.materialize().sum(dim, *args, **kwargs)

    def materialize(self) -> torch.Tensor:
        return self._parent.materialize()[:, self._level, :]

    def __getattr__(self, name):
        return getattr(self.materialize(), name)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        return _materialize_and_call(func, args, kwargs)


class _LevelSliceSum0:
    """prob[:, l, :].sum(0)  -- virtual (4, N); its .sum(0) is the level's column sum (N,)."""

    def __init__(self, parent: "LazyProbs", level: int):
        self._parent, self._level = parent, level
        self.shape = torch.Size(parent.shape[2:])

    def sum(self, dim=None, *args, **kwargs):
        if dim == 0 and not args and not kwargs:
            return self._parent.colsum[self._level]
        return self.materialize().sum(dim, *args, **kwargs)

    def materialize(self) -> torch.Tensor:
        return self._parent.materialize()[:, self._level, :].sum(0)

    def __getattr__(self, name):
        return getattr(self.materialize(), name)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        return _materialize_and_call(func, args, kwargs)


def _materialize_and_call(func, args, kwargs):
    conv = lambda a: a.materialize() if isinstance(a, (LazyProbs, _LevelSlice, _LevelSliceSum0)) else a
    args = tuple(conv(a) for a in args)
    kwargs = {k: conv(v) for k, v in (kwargs or {}).items()}
    return func(*args, **kwargs)


class LazyProbs:
    """Virtual (P, L, 4, N) tensor: row (p, l, v) == uvals[node(p, l, v)]."""

    def __init__(self, x: torch.Tensor, lat, uvals: torch.Tensor, colsum: torch.Tensor):
        self._x, self._lat, self._uvals = x, lat, uvals
        self.colsum = colsum                       # (L, N), differentiable
        self.shape = torch.Size((x.shape[0], lat.num_levels, 4, uvals.shape[1]))
        self.dtype, self.device = uvals.dtype, uvals.device
        self._dense = None

    # -- the uses the reference makes ------------------------------------------------------------------
    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def dim(self):
        return 4

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, key):
        if (isinstance(key, tuple) and len(key) == 3 and key[0] == slice(None) and isinstance(key[1], int)
                and key[2] == slice(None)):
            level = key[1] if key[1] >= 0 else key[1] + self.shape[1]
            return _LevelSlice(self, level)
        return self.materialize()[key]

    def clone(self):
        return self

    # -- everything else ---------------------------------------------------------------------------------
    def materialize(self) -> torch.Tensor:
        """The real (P, L, 4, N) tensor (differentiable gather of the per-node rows)."""
        if self._dense is None:
            self._dense = ops.GatherRows.apply(self._uvals, self._x, self._lat)
        return self._dense

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        return _materialize_and_call(func, args, kwargs)
