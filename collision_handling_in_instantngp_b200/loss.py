"""The reference's loss assembly (utils.py:78-174 + functions.py:243-245) evaluated from the column sums the
CUDA path produces, vectorised over levels.  `train_step` keeps using the reference's own ``Loss`` module
unchanged (it works on LazyProbs); this restatement exists so that bench.py and the data-parallel helper can
drive a step without importing the reference, and so that the (L, N) column sums can be all-reduced before
the non-linear divergence terms (dp.py).  Only O(L*N) torch ops -- no per-row tensor is ever formed."""
from __future__ import annotations

import math

import torch


def level_divergences(pbar: torch.Tensor, gamma: float, epsilon: float) -> torch.Tensor:
    """pbar (L,N) mean slot distribution -> (L,) of  -(gamma+epsilon) * JS + epsilon * KL  (utils.py:122-174).

    KLDivLoss('batchmean') on a 1-D vector divides by N (utils.py:86,144):
      kl = sum q (ln q - ln pbar) / N,  js = [sum m (ln m - ln pbar) + sum m (ln m - ln q)] / (2N), m=(pbar+q)/2
    """
    N = pbar.shape[-1]
    q = 1.0 / N
    lp = pbar.log()
    lq = math.log(q)
    kl = (q * (lq - lp)).sum(-1) / N
    m = (pbar + q) / 2
    lm = m.log()
    js = ((m * (lm - lp)).sum(-1) / N + (m * (lm - lq)).sum(-1) / N) / 2
    return -(gamma + epsilon) * js + epsilon * kl


def total_loss(rgb, target, colsum, rows: int, gamma: float, epsilon: float, l_mse: float = 1.0, l_js_kl: float = 1.0,
               collisions_term=None):
    """functions.py:243-245.  colsum (L,N) = sum over the `rows` = 4*P rows of each level of the probs tensor;
    collisions_term (L,) = l_collisions * collisions / (min_possible + delta), or None on epoch 0 (the scalar 1)."""
    mse = torch.nn.functional.mse_loss(rgb, target)
    levels = level_divergences(colsum / rows, gamma, epsilon)
    coll = torch.ones_like(levels) if collisions_term is None else collisions_term
    return l_mse * mse + (l_js_kl * levels + coll).sum(0), mse, levels
