"""Host-side geometry of the lattice on which the HPD is evaluated.

The reference feeds the HPD one row per (point, level, corner) (models.py:416-418), but the rows are integer
grid-corner coordinates and the HPD is shared by all levels, so only the distinct corners matter.  For a
batch whose coordinates lie in [lo, hi] the corners of level l are exactly the integer box
    floor(fl32(lo * n_l)) ... floor(fl32(hi * n_l)) + 1            (models.py:492-500, fp32 multiply)
and the union over levels is a box as well.  This module computes those boxes (same fp32 arithmetic as the
kernels) and fills the ``gngf_lattice`` struct of include/gngf.h.
"""
from __future__ import annotations

import numpy as np

from ._lib import MAX_LEVELS, Lattice


def level_resolutions(n_min: int, n_max: int, num_levels: int) -> np.ndarray:
    """n_l = floor(n_min * b**l) evaluated in numpy float64 exactly as models.py:305-317 does
    (n_max = 8192, L = 16 gives a finest level of 8191, not 8192)."""
    b = np.exp((np.log(n_max) - np.log(n_min)) / (num_levels - 1))
    return np.array([np.floor(n_min * b ** l) for l in range(num_levels)]).astype(np.int32)


def build_lattice(n_ls, lo=(0.0, 0.0), hi=(1.0, 1.0)) -> Lattice:
    n_ls = [int(n) for n in n_ls]
    L = len(n_ls)
    if not 0 < L <= MAX_LEVELS:
        raise ValueError(f"num_levels must be in 1..{MAX_LEVELS}, got {L}")
    lat = Lattice()
    lat.num_levels = L
    lo32 = np.asarray(lo, dtype=np.float32)
    hi32 = np.asarray(hi, dtype=np.float32)
    if not (np.isfinite(lo32).all() and np.isfinite(hi32).all() and (lo32 <= hi32).all()):
        raise ValueError(f"invalid coordinate bounds {lo} .. {hi}")
    off = 0
    gx0 = gy0 = 2 ** 31 - 1
    gx1 = gy1 = -2 ** 31
    for l, n in enumerate(n_ls):
        nf = np.float32(n)
        c0 = np.floor(lo32 * nf).astype(np.int64)          # floor corner of the smallest coordinate
        c1 = np.floor(hi32 * nf).astype(np.int64) + 1      # "+1" corner of the largest coordinate
        if n < 0:
            c0, c1 = np.minimum(c0, c1 - 1), np.maximum(c0 + 1, c1)
        lat.n[l] = n
        lat.lox[l], lat.loy[l] = int(c0[0]), int(c0[1])
        lat.lwx[l], lat.lwy[l] = int(c1[0] - c0[0] + 1), int(c1[1] - c0[1] + 1)
        lat.loff[l] = off
        off += lat.lwx[l] * lat.lwy[l]
        gx0, gy0 = min(gx0, int(c0[0])), min(gy0, int(c0[1]))
        gx1, gy1 = max(gx1, int(c1[0])), max(gy1, int(c1[1]))
    lat.loff[L] = off
    lat.ox, lat.oy = gx0, gy0
    lat.wx, lat.wy = gx1 - gx0 + 1, gy1 - gy0 + 1
    if lat.wx * lat.wy >= 2 ** 31 or off >= 2 ** 40:
        raise ValueError("lattice too large")
    return lat
