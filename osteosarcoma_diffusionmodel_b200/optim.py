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
        self._plans = {}
        self.last_grad_norm: Optional[torch.Tensor] = None      # device scalar: total gradient norm before clipping (last group stepped)

    def __del__(self):
        try:
            self._drop_plans()
        except Exception:
            pass

    def _plan(self, gi, group, ps):
        """Per-group launch plan, rebuilt only when the set of parameters with gradients (or their storage) changes: validated tensors,
        state tensors, the three pointer tables that do not change from step to step, the library handle."""
        key = tuple((id(p), p.data_ptr()) for p in ps)
        plan = self._plans.get(gi)
        if plan is not None and plan["key"] == key:
            return plan
        lib = _lib.load()
        for p in ps:
            if p.device.type != "cuda" or p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("FusedAdamW needs contiguous fp32 CUDA parameters; there is no CPU fallback")
            st = self.state[p]
            if not st:
                st["step"] = torch.tensor(0.0)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        steps = {int(self.state[p]["step"]) for p in ps}
        if len(steps) != 1:
            raise RuntimeError("FusedAdamW steps all parameters of a group together")
        # one shared step tensor per group (52 separate `+= 1` on CPU tensors cost 0.25 ms per step); state_dict() still shows one per parameter
        shared = torch.tensor(float(steps.pop()))
        for p in ps:
            self.state[p]["step"] = shared
        if plan is not None:
            lib.osteo_adamw_destroy(plan["handle"])
        h = C.c_void_p()
        n = len(ps)
        with torch.cuda.device(ps[0].device):
            _lib.check(lib.osteo_adamw_create(C.byref(h), n, (C.c_longlong * n)(*[p.numel() for p in ps])))
        ptr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])      # noqa: E731
        plan = dict(key=key, handle=h, n=n, step=shared, count=int(shared), device=ps[0].device, params=ptr(ps),
                    exp_avg=ptr([self.state[p]["exp_avg"] for p in ps]), exp_avg_sq=ptr([self.state[p]["exp_avg_sq"] for p in ps]),
                    grads=(C.c_void_p * n)(), norm=torch.empty((), device=ps[0].device, dtype=torch.float32))
        self._plans[gi] = plan
        return plan

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._drop_plans()

    def _drop_plans(self):
        lib = _lib.load()
        for plan in self._plans.values():
            lib.osteo_adamw_destroy(plan["handle"])
        self._plans = {}

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
            plan = self._plan(gi, group, ps)
            gtab = plan["grads"]
            for i, p in enumerate(ps):
                g = p.grad
                if g.dtype != torch.float32 or not g.is_contiguous():
                    raise RuntimeError("FusedAdamW needs contiguous fp32 gradients")
                gtab[i] = g.data_ptr()
            mg = group.get("max_grad_norm")
            b1, b2 = group["betas"]
            plan["count"] += 1
            with torch.cuda.device(plan["device"]):
                _lib.check(lib.osteo_adamw_step(plan["handle"], plan["params"], gtab, plan["exp_avg"], plan["exp_avg_sq"], float(group["lr"]), float(b1), float(b2),
                                                float(group["eps"]), float(group["weight_decay"]), plan["count"], float(mg) if mg else 0.0,
                                                plan["norm"].data_ptr() if mg else None, _lib.stream_handle()))
            plan["step"] += 1
            # the library wrote the parameters behind autograd's back: bump the version counters (the model's weight cache, like autograd's
            # saved-tensor checks, keys on them)
            _bump_versions(ps)
            self.last_grad_norm = plan["norm"] if mg else None
        return loss
