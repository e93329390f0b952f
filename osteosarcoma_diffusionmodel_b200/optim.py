"""Fused optimiser step for the training loop of utils/train.py:169-173, :242-244.

The reference runs `torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)` and `torch.optim.AdamW.step()` (lr 1e-4, weight decay
1e-5): with 52 parameter tensors torch's foreach implementations cost ~25 launches and ~0.9 ms of host time per step, which bounds
the whole training step once forward + backward are one graph launch.  `FusedAdamW` is a `torch.optim.Optimizer`
with AdamW's constructor, state layout (`step`, `exp_avg`, `exp_avg_sq`: optimizer checkpoints written by torch's AdamW load) and
arithmetic, whose `step()` is two kernels of the C-ABI library (`osteo_adamw_step`); with `max_grad_norm` set it also does the
clipping, so the `clip_grad_norm_` call can be dropped (calling it anyway is harmless: the second clip is the identity).
There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional

import torch

from . import _lib


def _bump_versions(tensors) -> None:
    setter = getattr(torch._C._autograd, "_unsafe_set_version_counter", None)
    if setter is not None:
        try:
            setter(list(tensors), [t._version + 1 for t in tensors])
            return
        except TypeError:
            pass
    torch._foreach_add_(list(tensors), 0.0)      # portable: an in-place no-op per tensor list


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-2,
                 max_grad_norm: Optional[float] = None):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError("invalid AdamW hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm))
        self._handles = {}
        self.last_grad_norm: Optional[torch.Tensor] = None      # device scalar: total gradient norm before clipping (last group stepped)

    def __del__(self):
        try:
            lib = _lib.load()
            for h, _ in self._handles.values():
                lib.osteo_adamw_destroy(h)
        except Exception:
            pass

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            for p in ps:
                if p.device.type != "cuda" or p.dtype != torch.float32 or not p.is_contiguous() or p.grad.dtype != torch.float32:
                    raise RuntimeError("FusedAdamW needs contiguous fp32 CUDA parameters and gradients; there is no CPU fallback")
                if not p.grad.is_contiguous():
                    p.grad = p.grad.contiguous()
                st = self.state[p]
                if not st:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            steps = {int(self.state[p]["step"]) for p in ps}
            if len(steps) != 1:
                raise RuntimeError("FusedAdamW steps all parameters of a group together")
            step = steps.pop() + 1
            sig = tuple(p.numel() for p in ps)
            key = (gi, ps[0].device)
            ent = self._handles.get(key)
            if ent is None or ent[1] != sig:
                if ent is not None:
                    lib.osteo_adamw_destroy(ent[0])
                h = C.c_void_p()
                arr = (C.c_longlong * len(ps))(*sig)
                with torch.cuda.device(ps[0].device):
                    _lib.check(lib.osteo_adamw_create(C.byref(h), len(ps), arr))
                ent = self._handles[key] = (h, sig)
            n = len(ps)
            tabs = [(C.c_void_p * n)(*[t.data_ptr() for t in ts]) for ts in
                    (ps, [p.grad for p in ps], [self.state[p]["exp_avg"] for p in ps], [self.state[p]["exp_avg_sq"] for p in ps])]
            mg = group.get("max_grad_norm")
            norm = torch.empty((), device=ps[0].device, dtype=torch.float32) if mg else None
            b1, b2 = group["betas"]
            with torch.cuda.device(ps[0].device):
                _lib.check(lib.osteo_adamw_step(ent[0], *tabs, float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]),
                                                step, float(mg) if mg else 0.0, _lib.ptr(norm), _lib.stream_handle()))
            for p in ps:
                self.state[p]["step"] += 1
            # the library wrote the parameters behind autograd's back: bump the version counters (the model's weight cache, like autograd's
            # saved-tensor checks, keys on them)
            _bump_versions(ps)
            self.last_grad_norm = norm
        return loss
