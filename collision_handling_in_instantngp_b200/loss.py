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


class _FusedLoss(torch.autograd.Function):
    """gngf_loss_fwd_bwd: forward value and both adjoints from one kernel."""

    @staticmethod
    def forward(ctx, rgb, target, colsum, rows, gamma, epsilon, l_mse, l_js_kl, coll_term):
        from . import ops
        rgb_c, tgt_c, cs_c = ops._f32c(rgb), ops._f32c(target), ops._f32c(colsum)
        L, N = cs_c.shape
        out = torch.empty(2 + L, dtype=torch.float32, device=rgb.device)
        d_rgb = torch.empty_like(rgb_c)
        d_colsum = torch.empty_like(cs_c)
        ops.call("gngf_loss_fwd_bwd", rgb_c.data_ptr(), tgt_c.data_ptr(), rgb_c.numel(), cs_c.data_ptr(), L, N,
                 float(rows), float(gamma), float(epsilon), float(l_mse), float(l_js_kl),
                 None if coll_term is None else ops._f32c(coll_term).data_ptr(), out.data_ptr(), d_rgb.data_ptr(),
                 d_colsum.data_ptr(), ops._stream())
        ctx.save_for_backward(d_rgb, d_colsum)
        total, mse, levels = out[0], out[1], out[2:]
        ctx.mark_non_differentiable(mse, levels)
        return total, mse, levels

    @staticmethod
    def backward(ctx, g_total, g_mse, g_levels):
        d_rgb, d_colsum = ctx.saved_tensors
        return d_rgb * g_total, None, d_colsum * g_total, None, None, None, None, None, None


def fused_total_loss(rgb, target, colsum, rows, gamma, epsilon, l_mse=1.0, l_js_kl=1.0, collisions_term=None):
    """Same value and gradients as :func:`total_loss`, from the single CUDA kernel of k7_loss.cu."""
    return _FusedLoss.apply(rgb, target, colsum, rows, gamma, epsilon, l_mse, l_js_kl, collisions_term)


def fused_loss_and_grads(rgb, target, colsum, rows, gamma, epsilon, l_mse=1.0, l_js_kl=1.0, collisions_term=None):
    """The same kernel without an autograd node: returns (out, d_rgb, d_colsum) with out = [total, mse, level_0..]
    and the adjoints of `total` w.r.t. rgb and colsum.  A caller whose loss is the root of the backward pass feeds
    them straight into ``torch.autograd.backward([rgb, colsum], [d_rgb, d_colsum])`` -- no ones_like(), no
    multiplications by the incoming gradient, no materialised zero gradients (trainer.GraphedTrainer)."""
    from . import ops
    rgb_c, tgt_c, cs_c = ops._f32c(rgb.detach()), ops._f32c(target), ops._f32c(colsum.detach())
    L, N = cs_c.shape
    out = torch.empty(2 + L, dtype=torch.float32, device=rgb_c.device)
    d_rgb = torch.empty_like(rgb_c)
    d_colsum = torch.empty_like(cs_c)
    ops.call("gngf_loss_fwd_bwd", rgb_c.data_ptr(), tgt_c.data_ptr(), rgb_c.numel(), cs_c.data_ptr(), L, N,
             float(rows), float(gamma), float(epsilon), float(l_mse), float(l_js_kl),
             None if collisions_term is None else ops._f32c(collisions_term).data_ptr(), out.data_ptr(),
             d_rgb.data_ptr(), d_colsum.data_ptr(), ops._stream())
    return out, d_rgb, d_colsum


def fused_loss_and_grads_split(rgb, target, colsum, rows, gamma, epsilon, l_mse=1.0, l_js_kl=1.0, collisions_term=None,
                               levels_stream=None):
    """As :func:`fused_loss_and_grads`, as two launches of the same kernel (gngf_loss_parts): the MSE half on the current
    stream, the divergence half on `levels_stream` -- the side stream on which the column sums are produced (and, under
    data parallelism, all-reduced) when the forward ran with ``ops.DEFER_COLSUM_JOIN``.  The caller joins that stream
    before it consumes `out` or `d_colsum` (GNGFPath.backward does, after the encoding's point pass).  All buffers are
    allocated on the current stream."""
    from . import ops
    rgb_c, tgt_c, cs_c = ops._f32c(rgb.detach()), ops._f32c(target), ops._f32c(colsum.detach())
    L, N = cs_c.shape
    out = torch.zeros(2 + L, dtype=torch.float32, device=rgb_c.device)
    d_rgb = torch.empty_like(rgb_c)
    d_colsum = torch.empty_like(cs_c)
    coll = None if collisions_term is None else ops._f32c(collisions_term)
    args = (rgb_c.data_ptr(), tgt_c.data_ptr(), rgb_c.numel(), cs_c.data_ptr(), L, N, float(rows), float(gamma),
            float(epsilon), float(l_mse), float(l_js_kl), None if coll is None else coll.data_ptr(), out.data_ptr(),
            d_rgb.data_ptr(), d_colsum.data_ptr())
    if levels_stream is None:
        ops.call("gngf_loss_parts", *args, 3, ops._stream())
        return out, d_rgb, d_colsum
    main = torch.cuda.current_stream()
    levels_stream.wait_stream(main)            # `out` is zeroed on the current stream
    ops.call("gngf_loss_parts", *args, 1, ops._stream())
    with torch.cuda.stream(levels_stream):
        ops.call("gngf_loss_parts", *args, 2, ops._stream())
    return out, d_rgb, d_colsum
