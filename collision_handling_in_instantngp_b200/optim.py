"""Fused Adam for the GNGF parameter set: the optimizer step of functions.py:96-127, 281 --
``torch.optim.Adam(groups, betas=(0.9, 0.99), eps=1e-15)`` with per-group ``lr`` / ``weight_decay`` -- as ONE kernel
launch over all parameter tensors (k9_adam.cu), with the step counter on the device so that the step can be captured
in a CUDA graph.  Same constructor shape as ``torch.optim.Adam`` (parameter groups with ``lr`` and ``weight_decay``;
``betas``; ``eps``); ``amsgrad`` / ``maximize`` are not supported (the reference does not use them).

The state dict is interchangeable with ``torch.optim.Adam``'s (the reference saves it as ``whole_opt.pt``,
functions.py:768): per parameter ``exp_avg``, ``exp_avg_sq`` and ``step``, the latter a float32 scalar tensor as in torch.
A state loaded from a stock Adam checkpoint (whose ``step`` lives on the CPU unless capturable / fused) is moved to the
parameter's device on the first step; a FusedAdam checkpoint loads into ``torch.optim.Adam`` as is.

    opt = FusedAdam([{"params": net.encoding.parameters(), "lr": 1e-4, "weight_decay": 0.0},
                     {"params": net.HPD.parameters(), "lr": 1e-3, "weight_decay": 1e-6},
                     {"params": net.mlp.parameters(), "lr": 1e-3, "weight_decay": 1e-6}], betas=(0.9, 0.99), eps=1e-15)
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import ADAM_MAX_TENSORS, AdamTensor, GngfError, check


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= weight_decay:
            raise ValueError("lr, eps and weight_decay must be non-negative")
        if not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError(f"invalid betas {betas}")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        betas_set = {tuple(g["betas"]) for g in self.param_groups}
        eps_set = {float(g["eps"]) for g in self.param_groups}
        if len(betas_set) != 1 or len(eps_set) != 1:
            raise GngfError("FusedAdam: betas and eps must be the same for all parameter groups")
        self._ticket = None

    def step_count(self, p) -> int:
        """Number of updates parameter `p` has received (device counter; this call synchronises)."""
        st = self.state.get(p)
        return 0 if not st else int(st["step"].item())

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        entries = []
        dev = None
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise GngfError("FusedAdam needs contiguous float32 CUDA parameters (there is no CPU path)")
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    g = g.float().contiguous()
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["step"] = torch.zeros((), dtype=torch.float32, device=p.device)   # advanced by the kernel
                else:
                    t = st["step"]      # a state loaded from torch.optim.Adam: CPU float scalar (or a Python number)
                    if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.float32 and t.device == p.device):
                        st["step"] = torch.as_tensor(float(t), dtype=torch.float32, device=p.device).reshape(())
                    for key in ("exp_avg", "exp_avg_sq"):
                        if st[key].device != p.device or st[key].dtype != torch.float32 or not st[key].is_contiguous():
                            st[key] = st[key].to(device=p.device, dtype=torch.float32).contiguous()
                dev = p.device
                entries.append((p, g, st["exp_avg"], st["exp_avg_sq"], st["step"], float(group["lr"]),
                                float(group["weight_decay"])))
        if not entries:
            return loss
        if self._ticket is None:
            self._ticket = torch.zeros(1, dtype=torch.int32, device=dev)
        beta1, beta2 = self.param_groups[0]["betas"]
        eps = float(self.param_groups[0]["eps"])
        stream = torch.cuda.current_stream(dev).cuda_stream
        fn = _lib.load().gngf_adam_step
        for i0 in range(0, len(entries), ADAM_MAX_TENSORS):     # (more than 64 tensors: several launches)
            chunk = entries[i0:i0 + ADAM_MAX_TENSORS]
            arr = (AdamTensor * len(chunk))()
            for i, (p, g, m, v, t, lr, wd) in enumerate(chunk):
                arr[i].p, arr[i].g, arr[i].m, arr[i].v = p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr()
                arr[i].step, arr[i].n, arr[i].lr, arr[i].weight_decay = t.data_ptr(), p.numel(), lr, wd
            check(fn(arr, len(chunk), float(beta1), float(beta2), eps, self._ticket.data_ptr(), stream), "gngf_adam_step")
        return loss
