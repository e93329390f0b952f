"""Drop-in `BiologyAwareDiffusionModel` whose compute runs on hand-written sm_100a kernels.

Mirrors the reference class contract (models/diffusion.py:259-449): same
constructor, same attributes, the same 52 parameters + 4 buffers under the same state_dict keys,
`forward(x_0, conditions, return_loss=True)`, `q_sample`, `p_sample`, `sample`.  The nn.Module tree
below only CONTAINS the parameters (so `.to()`, `.parameters()`, `load_state_dict(strict=True)`,
AdamW and `clip_grad_norm_` behave exactly as with the reference); every forward / backward /
sampling computation is a call into the C-ABI library (include/osteo_ddpm.h).  There is no eager
PyTorch implementation here and no CPU fallback: calling a compute method on a CPU model raises.
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import _lib

__all__ = ["BiologyAwareDiffusionModel", "ConditionalEmbedding", "TimeEmbedding", "DiffusionUNet"]

_PRECISIONS = {"bf16": _lib.PREC_BF16, "fp32x3": _lib.PREC_FP32X3}
_ALWAYS_REPACK = bool(int(os.environ.get("OSTEO_DDPM_ALWAYS_REPACK", "0")))


class ConditionalEmbedding(nn.Module):
    """Parameter container for models/diffusion.py:91-114 (Linear C->E, SiLU, Linear E->E)."""

    def __init__(self, num_continuous: int, embedding_dim: int):
        super().__init__()
        self.num_continuous = num_continuous
        self.embedding_dim = embedding_dim
        self.mlp = nn.Sequential(nn.Linear(num_continuous, embedding_dim), nn.SiLU(), nn.Linear(embedding_dim, embedding_dim))


class TimeEmbedding(nn.Module):
    """models/diffusion.py:117-139.  Only used on the host to tabulate the embedding of the
    num_steps integer timesteps once (the kernels gather rows of the projected table)."""

    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim

    def table(self, num_steps: int) -> torch.Tensor:
        half = self.dim // 2
        e = np.log(10000) / (half - 1)
        e = torch.exp(torch.arange(half) * -e)
        # p_sample: torch.full((B,), t / T) (models/diffusion.py:392); forward: t.float() / T (:367).
        t_norm = torch.tensor([t / num_steps for t in range(num_steps)], dtype=torch.float32)
        e = t_norm[:, None] * e[None, :]
        return torch.cat([torch.sin(e), torch.cos(e)], dim=-1).contiguous()


class DiffusionUNet(nn.Module):
    """Parameter container with the module tree of models/diffusion.py:142-196."""

    def __init__(self, data_dim: int, time_dim: int = 128, condition_dim: int = 64, hidden_dims: Sequence[int] = (256, 512, 256), dropout: float = 0.1):
        super().__init__()
        hidden_dims = list(hidden_dims)
        self.data_dim = data_dim
        self.hidden_dims = hidden_dims
        self.time_embed = TimeEmbedding(time_dim)
        self.input_proj = nn.Linear(data_dim, hidden_dims[0])
        self.cond_proj = nn.Linear(condition_dim, hidden_dims[0])
        self.time_proj = nn.Linear(time_dim, hidden_dims[0])
        self.encoder = nn.ModuleList()
        in_dim = hidden_dims[0]
        for h_dim in hidden_dims[1:]:
            self.encoder.append(self._make_block(in_dim, h_dim, dropout))
            in_dim = h_dim
        self.bottleneck = self._make_block(in_dim, in_dim, dropout)
        self.decoder = nn.ModuleList()
        curr = hidden_dims[-1]
        for i in range(len(hidden_dims) - 2, -1, -1):
            self.decoder.append(self._make_block(curr + hidden_dims[i + 1], hidden_dims[i], dropout))
            curr = hidden_dims[i]
        self.output_proj = nn.Linear(curr, data_dim)

    @staticmethod
    def _make_block(in_dim: int, out_dim: int, dropout: float) -> nn.Sequential:
        return nn.Sequential(nn.Linear(in_dim, out_dim), nn.GroupNorm(8, out_dim), nn.SiLU(), nn.Dropout(dropout),
                             nn.Linear(out_dim, out_dim), nn.GroupNorm(8, out_dim), nn.SiLU())

    def blocks(self) -> List[nn.Sequential]:
        return list(self.encoder) + [self.bottleneck] + list(self.decoder)


class _TrainStep(torch.autograd.Function):
    """loss = model(x0, cond) with the whole forward+backward fused into one C-ABI call; autograd only
    scales the precomputed parameter gradients by the incoming d(loss)."""

    @staticmethod
    def forward(ctx, model, x0, cond, inject, *params):
        ws = 1
        if model._flat_allreduce:
            from . import distributed as D
            ws = D.world()[1]
        # data parallel: the flat gradient buffer (17 MB for config.yaml) is all-reduced in place instead of per-bucket cat / copy-back --
        # in two pieces with the backward pass cut between them when model._dp_overlap (the tail's all-reduce runs beside the second
        # part of the backward), else as ONE call after the whole step; the 1 / world_size of the mean rides on the scaling backward()
        # does anyway
        overlap = ws > 1 and model._dp_overlap
        loss, grads = model._run_train_step(x0, cond, inject, want_grads=True, dp_overlap=overlap)
        ctx.model, ctx.grads, ctx.epoch = model, grads, model._grad_epoch
        ctx.scale = 1.0
        if ws > 1:
            if not overlap:
                D.all_reduce_sum_(model._grad_buf[0])
            ctx.scale = 1.0 / ws
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        # The gradients live in the model's persistent flat buffer (stable addresses: the library replays the step as one graph);
        # a second forward before this backward would have overwritten them.
        if ctx.model._grad_epoch != ctx.epoch:
            raise RuntimeError("backward() of a loss whose gradients were overwritten by a later forward(): call backward() before the next "
                               "training forward of this model")
        # one fused multi-tensor kernel instead of 52 tiny launches; the products are fresh tensors
        return (None, None, None, None) + tuple(torch._foreach_mul(ctx.grads, grad_out * ctx.scale if ctx.scale != 1.0 else grad_out))


def _model_device(model):
    try:
        return model._param_list()[0].device
    except Exception:       # half-constructed / torn-down module
        return None


_on_model_device = _lib.on_device(_model_device)


class BiologyAwareDiffusionModel(nn.Module):
    """B200-native mirror of models/diffusion.py:259-449."""

    def __init__(self, mutation_dim: int, expression_dim: int, pathway_dim: int, condition_dim: int, config: dict):
        super().__init__()
        self.mutation_dim = mutation_dim
        self.expression_dim = expression_dim
        self.pathway_dim = pathway_dim
        self.condition_dim = condition_dim
        self.data_dim = mutation_dim + expression_dim + pathway_dim

        mcfg = config["model"]
        self._latent_dim = int(mcfg["latent_dim"])
        self._hidden_dims = [int(h) for h in mcfg["hidden_dims"]]
        self._dropout = float(mcfg["gnn"]["dropout"])
        self.condition_embed = ConditionalEmbedding(num_continuous=condition_dim, embedding_dim=64)
        self.unet = DiffusionUNet(data_dim=self.data_dim, time_dim=self._latent_dim, condition_dim=self._latent_dim // 2,
                                  hidden_dims=self._hidden_dims, dropout=self._dropout)
        if self._latent_dim // 2 != 64:
            # the reference hard-codes embedding_dim=64 (models/diffusion.py:285) against latent_dim // 2 (:292)
            raise ValueError("latent_dim must be 128: ConditionalEmbedding emits 64 features (models/diffusion.py:285,292)")

        self.num_steps = int(mcfg["diffusion"]["num_steps"])
        self.register_buffer("betas", self._get_beta_schedule(mcfg["diffusion"]["beta_schedule"], self.num_steps))
        alphas = 1.0 - self.betas
        alphas_cumprod = torch.cumprod(alphas, dim=0)
        self.register_buffer("alphas_cumprod", alphas_cumprod)
        self.register_buffer("sqrt_alphas_cumprod", torch.sqrt(alphas_cumprod))
        self.register_buffer("sqrt_one_minus_alphas_cumprod", torch.sqrt(1.0 - alphas_cumprod))

        b200 = mcfg.get("b200", {}) if isinstance(mcfg.get("b200", {}), dict) else {}
        self._precision = str(os.environ.get("OSTEO_DDPM_PRECISION", b200.get("precision", "bf16")))
        if self._precision not in _PRECISIONS:
            raise ValueError(f"unknown precision {self._precision!r} (bf16 | fp32x3)")
        self._chunk_rows = int(os.environ.get("OSTEO_DDPM_CHUNK_ROWS", b200.get("chunk_rows", 131072)))
        self._use_graph = bool(int(os.environ.get("OSTEO_DDPM_GRAPH", b200.get("use_graph", 1))))
        self._fused = bool(int(os.environ.get("OSTEO_DDPM_FUSED", b200.get("fused", 1))))
        self._branches = b200.get("branches")          # None = library default
        self._train_graph = bool(int(os.environ.get("OSTEO_TRAIN_GRAPH", b200.get("train_graph", 1))))
        self._seed = int(b200.get("seed", 0))
        self._draws = 0                 # counter mixed into the seed of un-seeded calls
        self._grad_buf = None           # (flat fp32 gradient buffer, per-parameter views, ctypes pointer array)
        self._grad_epoch = 0            # bumped by every gradient-producing forward
        self._flat_allreduce = False    # distributed.dp_train_step: average the flat gradient buffer over the ranks inside forward()
        # ... in two pieces, the first beside the second part of the backward pass. Opt-in: measured 1.33 vs 1.28 ms per step on two GPUs
        # (three graph launches + two async NCCL calls cost more than the ~50 us of all-reduce they hide)
        self._dp_overlap = os.environ.get("OSTEO_DP_OVERLAP", "0") == "1"
        self._ctx = None                # C context handle
        self._ctx_device = None
        self._weights_sig = None
        self._schedule_sig = None
        self._train_enabled = False
        self._inject: Optional[dict] = None   # test hook: {"t":..., "noise":..., "masks":[...]} consumed by forward()

    # ------------------------------------------------------------------ schedule (models/diffusion.py:312-326)
    def _get_beta_schedule(self, schedule_type: str, num_steps: int) -> torch.Tensor:
        if schedule_type == "linear":
            return torch.linspace(1e-4, 0.02, num_steps)
        elif schedule_type == "cosine":
            steps = torch.arange(num_steps + 1, dtype=torch.float32) / num_steps
            alphas_cumprod = torch.cos((steps + 0.008) / 1.008 * np.pi / 2) ** 2
            alphas_cumprod = alphas_cumprod / alphas_cumprod[0]
            betas = 1 - (alphas_cumprod[1:] / alphas_cumprod[:-1])
            return torch.clip(betas, 0.0001, 0.9999)
        else:
            raise ValueError(f"Unknown schedule: {schedule_type}")

    @staticmethod
    def reverse_coefficients(betas: torch.Tensor, alphas_cumprod: torch.Tensor):
        """x_{t-1} = c_x[t]*x_t - c_eps[t]*eps + sigma[t]*z: the reference's x0_pred / two-term posterior mean /
        variance (models/diffusion.py:400-423) collapsed in fp64 from its own fp32 buffers."""
        b = betas.detach().double().cpu().numpy()
        ab = alphas_cumprod.detach().double().cpu().numpy()
        T = len(b)
        cx, ce, sg = np.zeros(T), np.zeros(T), np.zeros(T)
        for t in range(T):
            sa, s1 = math.sqrt(ab[t]), math.sqrt(1.0 - ab[t])
            if t == 0:
                cx[t], ce[t], sg[t] = 1.0 / sa, s1 / sa, 0.0
                continue
            abp = ab[t - 1]
            k0 = math.sqrt(abp) * b[t] / (1.0 - ab[t])
            k1 = math.sqrt(1.0 - b[t]) * (1.0 - abp) / (1.0 - ab[t])
            cx[t], ce[t] = k0 / sa + k1, k0 * s1 / sa
            sg[t] = math.sqrt((1.0 - abp) / (1.0 - ab[t]) * b[t])
        return cx, ce, sg

    # ------------------------------------------------------------------ configuration knobs
    @_on_model_device
    def set_precision(self, precision: str) -> "BiologyAwareDiffusionModel":
        """'bf16' (bf16 operands, fp32 accumulate) or 'fp32x3' (split-bf16, three tensor-core passes, ~fp32)."""
        if precision not in _PRECISIONS:
            raise ValueError(f"unknown precision {precision!r} (bf16 | fp32x3)")
        self._precision = precision
        if self._ctx is not None:
            _lib.check(_lib.load().osteo_ddpm_set_precision(self._ctx, _PRECISIONS[precision]))
            self._weights_sig = None
        return self

    @_on_model_device
    def set_fused(self, enable: bool) -> "BiologyAwareDiffusionModel":
        """bf16 sampling runs the fused step kernel (output_proj + reverse update + next input_proj) by default;
        set_fused(False) selects the unfused kernels (always used by 'fp32x3')."""
        self._fused = bool(enable)
        if self._ctx is not None:
            _lib.check(_lib.load().osteo_ddpm_set_fused(self._ctx, int(self._fused)))
        return self

    @_on_model_device
    def set_chunk_rows(self, rows: int) -> None:
        self._chunk_rows = int(rows)
        if self._ctx is not None:
            _lib.check(_lib.load().osteo_ddpm_set_chunk_rows(self._ctx, self._chunk_rows))

    @_on_model_device
    def set_train_graph(self, enable: bool) -> None:
        """Training steps with nothing injected are replayed as one executable graph per (batch size, gradient buffer) by default;
        set_train_graph(False) keeps every launch eager (same results)."""
        self._train_graph = bool(enable)
        if self._ctx is not None:
            _lib.check(_lib.load().osteo_ddpm_set_train_graph(self._ctx, int(self._train_graph)))

    @_on_model_device
    def set_branches(self, branches: int) -> None:
        """Number of parallel row branches of the sampling graphs (1..4, default 2); results do not depend on it."""
        self._branches = int(branches)
        if self._ctx is not None:
            _lib.check(_lib.load().osteo_ddpm_set_branches(self._ctx, self._branches))

    def manual_seed(self, seed: int) -> None:
        """Seed of the in-kernel Philox streams (x_T, reverse noise, q_sample noise, dropout, timesteps)."""
        self._seed = int(seed)
        self._draws = 0

    def _next_seed(self) -> int:
        self._draws += 1
        return (self._seed * 0x9E3779B97F4A7C15 + self._draws * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF

    # ------------------------------------------------------------------ C context management
    def _device(self) -> torch.device:
        return self._param_list()[0].device

    def _param_list(self) -> List[torch.Tensor]:
        """Parameters in the order osteo_ddpm_set_weights expects (state_dict order). Cached: walking the Sequential containers
        costs ~0.1 ms, several times per training step; Module._apply (.to / .cuda) keeps the Parameter objects, and the cache is
        dropped whenever a first or last entry has been replaced."""
        ce, u = self.condition_embed.mlp, self.unet
        ps = self.__dict__.get("_plist")
        if ps is not None and ps[0] is ce._modules["0"]._parameters["weight"] and ps[-1] is u.output_proj._parameters["bias"]:
            return ps
        ps = [ce[0].weight, ce[0].bias, ce[2].weight, ce[2].bias, u.input_proj.weight, u.input_proj.bias,
              u.cond_proj.weight, u.cond_proj.bias, u.time_proj.weight, u.time_proj.bias]
        for blk in u.blocks():
            ps += [blk[0].weight, blk[0].bias, blk[1].weight, blk[1].bias, blk[4].weight, blk[4].bias, blk[5].weight, blk[5].bias]
        ps += [u.output_proj.weight, u.output_proj.bias]
        self.__dict__["_plist"] = ps
        return ps

    def _ensure_ctx(self, rows: int, train: bool = False):
        dev = self._device()
        if dev.type != "cuda":
            raise RuntimeError("BiologyAwareDiffusionModel (B200-native) computes only on a CUDA device: move the model with "
                               ".to('cuda'); there is no CPU fallback")
        lib = _lib.load()
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        if self._ctx is None or self._ctx_device != index:
            self._destroy_ctx()
            handle = C.c_void_p()
            hid = (C.c_int * len(self._hidden_dims))(*self._hidden_dims)
            _lib.check(lib.osteo_ddpm_create(C.byref(handle), index, self.data_dim, self.condition_dim, self._latent_dim, 64,
                                             len(self._hidden_dims), hid, self.num_steps, self._dropout, _PRECISIONS[self._precision]))
            self._ctx, self._ctx_device = handle, index
            self._weights_sig = self._schedule_sig = None
            self._train_enabled = False
            _lib.check(lib.osteo_ddpm_set_chunk_rows(self._ctx, self._chunk_rows))
            _lib.check(lib.osteo_ddpm_set_fused(self._ctx, int(self._fused)))
            if self._branches is not None:
                _lib.check(lib.osteo_ddpm_set_branches(self._ctx, self._branches))
            _lib.check(lib.osteo_ddpm_set_train_graph(self._ctx, int(self._train_graph)))
            emb = self.unet.time_embed.table(self.num_steps).numpy()
            _lib.check(lib.osteo_ddpm_set_time_embedding(self._ctx, emb.ctypes.data))
        if rows > lib.osteo_ddpm_capacity(self._ctx):
            _lib.check(lib.osteo_ddpm_reserve(self._ctx, int(rows)))
        if train and not self._train_enabled:
            _lib.check(lib.osteo_ddpm_enable_training(self._ctx, 1))
            self._train_enabled = True
            self._weights_sig = None
        self._sync_schedule()
        self._sync_weights()
        return lib

    def _sync_schedule(self) -> None:
        sig = (self.betas.data_ptr(), self.betas._version, self.alphas_cumprod._version,
               self.sqrt_alphas_cumprod._version, self.sqrt_one_minus_alphas_cumprod._version)
        if sig == self._schedule_sig:
            return
        cx, ce, sg = self.reverse_coefficients(self.betas, self.alphas_cumprod)
        arrs = [self.sqrt_alphas_cumprod.detach().float().cpu().numpy(), self.sqrt_one_minus_alphas_cumprod.detach().float().cpu().numpy(),
                cx.astype(np.float32), ce.astype(np.float32), sg.astype(np.float32)]
        arrs = [np.ascontiguousarray(a) for a in arrs]
        _lib.check(_lib.load().osteo_ddpm_set_schedule(self._ctx, *[a.ctypes.data for a in arrs]))
        self._schedule_sig = sig

    def _sync_weights(self) -> None:
        ps = self._param_list()
        sig = tuple((p.data_ptr(), p._version) for p in ps) + (self._precision,)
        if sig == self._weights_sig and not _ALWAYS_REPACK:
            return
        for p in ps:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("parameters must be contiguous fp32 tensors (the reference is fp32-only, SURVEY.md §0.5)")
        arr = (C.c_void_p * len(ps))(*[p.data_ptr() for p in ps])
        _lib.check(_lib.load().osteo_ddpm_set_weights(self._ctx, arr, len(ps), _lib.stream_handle()))
        self._weights_sig = sig

    def invalidate_weights(self) -> None:
        """Force the library to repack the parameters before the next compute call. The repack is normally triggered by the
        parameters' (address, version) signature; in-place writes THROUGH `.data` (`p.data.copy_(ema)`, init code using `.data`)
        bump a separate version counter torch does not expose on the Parameter, so after such writes call this (load_state_dict,
        `.to()` / `.cuda()` / `.float()` and optimizer steps are detected automatically). Set OSTEO_DDPM_ALWAYS_REPACK=1 to repack
        before every call instead (slower: ~30 small launches)."""
        self._weights_sig = None
        self._schedule_sig = None

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_weights()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.__dict__["_plist"] = None
        self.invalidate_weights()
        return out

    def _destroy_ctx(self) -> None:
        ctx = self.__dict__.get("_ctx")
        if ctx is not None:
            self.__dict__["_ctx"] = None
            try:
                _lib.load().osteo_ddpm_destroy(ctx)
            except Exception:
                pass

    def __del__(self):
        try:
            self._destroy_ctx()
        except Exception:      # interpreter shutdown: module globals may already be gone
            pass

    def __getstate__(self):
        # the C context is per-object device state: copies / pickles start without one
        state = self.__dict__.copy()
        state.update(_ctx=None, _ctx_device=None, _weights_sig=None, _schedule_sig=None, _inject=None, _train_enabled=False, _grad_buf=None, _plist=None)
        return state

    @_on_model_device
    def check_status(self) -> None:
        """Raise if any kernel pipeline reported a (bounded-wait) timeout. Synchronises the current stream."""
        if self._ctx is not None:
            _lib.check(_lib.load().osteo_ddpm_status(self._ctx, _lib.stream_handle()))

    def sampling_mode(self) -> dict:
        """How the last sample() ran: {'precision', 'fused' (the fused step kernel), 'graph_branches' (row branches of the replayed graph,
        0 = launched eagerly)} -- lets tests and bench.py assert that the benchmarked configuration is the one that was exercised."""
        lib = _lib.load()
        if self._ctx is None:
            return {"precision": self._precision, "fused": False, "graph_branches": 0}
        return {"precision": self._precision, "fused": bool(lib.osteo_ddpm_step_is_fused(self._ctx)),
                "graph_branches": int(lib.osteo_ddpm_graph_branches(self._ctx)) if self._use_graph else 0}

    def launch_count(self) -> int:
        return int(_lib.load().osteo_ddpm_launch_count(self._ctx)) if self._ctx is not None else 0

    @staticmethod
    def _as_f32(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
        return t.to(device=dev, dtype=torch.float32).contiguous()

    # ------------------------------------------------------------------ forward process (models/diffusion.py:328-342)
    @torch.no_grad()
    @_on_model_device
    def q_sample(self, x_0, t, noise=None):
        dev = self._device()
        lib = self._ensure_ctx(1)
        x_0 = self._as_f32(x_0, dev)
        n = x_0.shape[0]
        t32 = t.to(device=dev, dtype=torch.int32).contiguous()
        gen = noise is None
        noise_t = torch.empty_like(x_0) if gen else self._as_f32(noise, dev)      # injected noise is only read (the reference returns the caller's tensor too)
        x_t = torch.empty_like(x_0)
        _lib.check(lib.osteo_ddpm_q_sample(self._ctx, x_0.data_ptr(), t32.data_ptr(), noise_t.data_ptr(), x_t.data_ptr(), n, int(gen),
                                           self._next_seed() if gen else 0, 0, 0, _lib.stream_handle()))
        return x_t, noise_t

    # ------------------------------------------------------------------ training forward (models/diffusion.py:344-380)
    @_on_model_device
    def forward(self, x_0, conditions, return_loss=True):
        dev = self._device()
        x_0 = self._as_f32(x_0, dev)
        conditions = self._as_f32(conditions, dev)
        inject = self._inject
        self._inject = None
        if return_loss:
            if torch.is_grad_enabled() and any(p.requires_grad for p in self._param_list()):
                return _TrainStep.apply(self, x_0, conditions, inject, *self._param_list())
            loss, _ = self._run_train_step(x_0, conditions, inject, want_grads=False)
            return loss
        # noise prediction only
        n = x_0.shape[0]
        lib = self._ensure_ctx(n)
        t = self._draw_t(n, inject)
        x_t, _ = self.q_sample(x_0, t, None if inject is None else inject.get("noise"))
        eps = torch.empty_like(x_0)
        s = _lib.stream_handle()
        _lib.check(lib.osteo_ddpm_set_conditions(self._ctx, conditions.data_ptr(), n, s))
        _lib.check(lib.osteo_ddpm_denoise(self._ctx, x_t.data_ptr(), t.to(torch.int32).contiguous().data_ptr(), n, eps.data_ptr(), s))
        return eps

    def _draw_t(self, n: int, inject: Optional[dict]) -> torch.Tensor:
        if inject is not None and inject.get("t") is not None:
            return inject["t"].to(self._device())
        return torch.randint(0, self.num_steps, (n,), device=self._device())   # models/diffusion.py:361

    def _dp_cut(self, ps) -> tuple:
        """(cut, offset): the backward pass is cut before half block `cut` so that the gradients it has finished by then -- the tensors
        [10 + 4 cut, end) of the state_dict order, elements [offset, total) of the flat buffer -- are about half of the bytes."""
        sizes = [p.numel() for p in ps]
        total, n_halves = sum(sizes), (len(ps) - 12) // 4
        best = (n_halves, sum(sizes[:10 + 4 * n_halves]))
        for cut in range(1, n_halves):
            off = sum(sizes[:10 + 4 * cut])
            if abs(off - total // 2) < abs(best[1] - total // 2):
                best = (cut, off)
        return best

    @_on_model_device
    def _run_train_step(self, x_0, conditions, inject, want_grads: bool, aux=None, dp_overlap: bool = False):
        """One C-ABI training step. `aux` (multitask.py) adds auxiliary losses on the predicted clean sample: the step then runs in
        two halves (osteo_ddpm_train_forward / _backward) with d(aux)/d(x0hat) injected between them."""
        n = x_0.shape[0]
        lib = self._ensure_ctx(n, train=want_grads)
        dev = self._device()
        t = self._draw_t(n, inject).to(torch.int32).contiguous()
        noise = None
        masks = None
        if inject is not None:
            if inject.get("noise") is not None:
                noise = self._as_f32(inject["noise"], dev)
            if inject.get("masks") is not None:
                masks = [m.to(device=dev, dtype=torch.uint8).contiguous() for m in inject["masks"]]
        loss = torch.zeros((), device=dev, dtype=torch.float32)
        ps = self._param_list()
        grads, garr = [], None
        if want_grads:
            # one persistent flat buffer, per-parameter views: the library zeroes it with a single memset, and the stable addresses
            # let it replay the whole step as one executable graph (osteo_ddpm_set_train_graph)
            total = sum(p.numel() for p in ps)
            gb = self._grad_buf
            if gb is None or gb[0].device != dev or gb[0].numel() != total:
                flat = torch.empty(total, device=dev, dtype=torch.float32)
                views, o = [], 0
                for p in ps:
                    views.append(flat[o:o + p.numel()].view_as(p))
                    o += p.numel()
                gb = self._grad_buf = (flat, views, (C.c_void_p * len(ps))(*[g.data_ptr() for g in views]))
            _, grads, garr = gb
            self._grad_epoch += 1
        marr = (C.c_void_p * len(masks))(*[m.data_ptr() for m in masks]) if masks is not None else None
        seed, s = self._next_seed(), _lib.stream_handle()
        if dp_overlap and aux is None and want_grads:
            # forward, then the backward pass in two launches with the tail's all-reduce started between them (NCCL waits for what is on the
            # current stream when it is called, i.e. for part 1 only, and runs on its own stream beside part 2)
            import torch.distributed as dist
            flat = self._grad_buf[0]
            cut, off = self._dp_cut(ps)
            _lib.check(lib.osteo_ddpm_train_forward(self._ctx, x_0.data_ptr(), conditions.data_ptr(), n, t.data_ptr(), _lib.ptr(noise), marr, int(self.training),
                                                    seed, 0, loss.data_ptr(), s))
            _lib.check(lib.osteo_ddpm_train_backward_part(self._ctx, conditions.data_ptr(), n, t.data_ptr(), marr, int(self.training), seed, 0, garr, len(ps), 1, cut, s))
            h_tail = dist.all_reduce(flat[off:], op=dist.ReduceOp.SUM, async_op=True)
            _lib.check(lib.osteo_ddpm_train_backward_part(self._ctx, conditions.data_ptr(), n, t.data_ptr(), marr, int(self.training), seed, 0, garr, len(ps), 2, cut, s))
            h_head = dist.all_reduce(flat[:off], op=dist.ReduceOp.SUM, async_op=True)
            h_tail.wait()
            h_head.wait()
            return loss, grads
        if aux is None or not want_grads:
            _lib.check(lib.osteo_ddpm_train_step(self._ctx, x_0.data_ptr(), conditions.data_ptr(), n, t.data_ptr(), _lib.ptr(noise), marr, int(self.training),
                                                 seed, 0, loss.data_ptr(), garr, len(ps) if want_grads else 0, s))
            return loss, grads
        _lib.check(lib.osteo_ddpm_train_forward(self._ctx, x_0.data_ptr(), conditions.data_ptr(), n, t.data_ptr(), _lib.ptr(noise), marr, int(self.training),
                                                seed, 0, loss.data_ptr(), s))
        cols = aux.columns.to(device=dev, dtype=torch.int32).contiguous()
        x0hat = torch.empty((n, cols.numel()), device=dev, dtype=torch.float32)
        _lib.check(lib.osteo_ddpm_train_x0hat(self._ctx, x_0.data_ptr(), t.data_ptr(), n, cols.data_ptr(), cols.numel(), x0hat.data_ptr(), s))
        g = aux.gradient(x0hat, t)                # d(aux loss)/d(x0hat[:, cols]) or None
        if g is not None:
            g = g.to(torch.float32).contiguous()
            _lib.check(lib.osteo_ddpm_train_inject(self._ctx, t.data_ptr(), n, cols.data_ptr(), cols.numel(), g.data_ptr(), s))
        _lib.check(lib.osteo_ddpm_train_backward(self._ctx, conditions.data_ptr(), n, t.data_ptr(), marr, int(self.training), seed, 0, garr, len(ps), s))
        return loss, grads

    # ------------------------------------------------------------------ reverse process (models/diffusion.py:382-449)
    @torch.no_grad()
    @_on_model_device
    def p_sample(self, x_t, t, conditions, noise=None, return_eps: bool = False, seed: Optional[int] = None):
        """Single reverse step.  `noise` injects z (parity); otherwise z comes from the in-kernel Philox stream."""
        dev = self._device()
        x_t = self._as_f32(x_t, dev)
        conditions = self._as_f32(conditions, dev)
        n = x_t.shape[0]
        lib = self._ensure_ctx(n)
        s = _lib.stream_handle()
        t = int(t)
        _lib.check(lib.osteo_ddpm_load_state(self._ctx, x_t.data_ptr(), n, s))
        _lib.check(lib.osteo_ddpm_set_conditions(self._ctx, conditions.data_ptr(), n, s))
        z = self._as_f32(noise, dev) if noise is not None else None
        eps = torch.empty_like(x_t) if return_eps else None
        _lib.check(lib.osteo_ddpm_reverse_step(self._ctx, n, t, _lib.ptr(z), _lib.ptr(eps), self._next_seed() if seed is None else seed, 0, s))
        out = torch.empty_like(x_t)
        _lib.check(lib.osteo_ddpm_store_state(self._ctx, out.data_ptr(), n, s))
        return (out, eps) if return_eps else out

    @torch.no_grad()
    @_on_model_device
    def sample(self, conditions, num_samples: int = 1, *, seed: Optional[int] = None, row_base: int = 0, x_T=None, noise=None,
               t_stop: int = 0, _components: Optional[str] = None):
        """Generate samples via reverse diffusion (models/diffusion.py:427-449). Sampling always runs the denoiser in inference form
        (no dropout), whatever `self.training` says: every reference caller samples under `.eval()` (utils/generate.py:29,
        models/diffusion.py:476), where the reference's Dropout is the identity too.

        Extra keyword-only arguments (all optional, defaults reproduce the reference call):
          seed / row_base  Philox key and global index of row 0: rows get the same noise however the cohort is
                           sharded over GPUs.
          x_T, noise       injected start state [n, D] and per-step z [T - t_stop, n, D] (parity runs).
          t_stop           stop after timestep t_stop (0 = full loop).
        """
        dev = self._device()
        conditions = self._as_f32(conditions, dev)
        n = int(num_samples)
        if conditions.shape[0] != n:
            raise ValueError(f"conditions has {conditions.shape[0]} rows but num_samples={n}")
        lib = self._ensure_ctx(n)
        s = _lib.stream_handle()
        seed = self._next_seed() if seed is None else int(seed)
        _lib.check(lib.osteo_ddpm_set_conditions(self._ctx, conditions.data_ptr(), n, s))
        if x_T is not None:
            x_T = self._as_f32(x_T, dev)
            _lib.check(lib.osteo_ddpm_load_state(self._ctx, x_T.data_ptr(), n, s))
        else:
            _lib.check(lib.osteo_ddpm_init_noise(self._ctx, n, seed, int(row_base), s))
        z = self._as_f32(noise, dev) if noise is not None else None
        _lib.check(lib.osteo_ddpm_sample_loop(self._ctx, n, self.num_steps - 1, int(t_stop), _lib.ptr(z), seed, int(row_base),
                                              int(self._use_graph), s))
        if _components:
            return self._store_components(n, dev, lib, s, conditions, pack_bits=_components == "bits")
        out = torch.empty((n, self.data_dim), device=dev, dtype=torch.float32)
        _lib.check(lib.osteo_ddpm_store_state(self._ctx, out.data_ptr(), n, s))
        return out

    def _store_components(self, n, dev, lib, s, conditions, pack_bits: bool):
        md = self.mutation_dim
        calls = torch.empty((n, md), device=dev, dtype=torch.uint8)
        bits = torch.empty((n, (md + 7) // 8), device=dev, dtype=torch.uint8) if pack_bits else None
        rest = torch.empty((n, self.data_dim - md), device=dev, dtype=torch.float32)
        _lib.check(lib.osteo_ddpm_store_split(self._ctx, n, md, 0.5, calls.data_ptr(), _lib.ptr(bits), rest.data_ptr(), s))
        out = {"mutations": calls, "expression": rest[:, :self.expression_dim], "pathways": rest[:, self.expression_dim:], "conditions": conditions}
        if pack_bits:
            out["mutation_bits"] = bits
        return out

    @torch.no_grad()
    @_on_model_device
    def sample_components(self, conditions, num_samples: int = 1, *, pack_bits: bool = False, **kw):
        """sample() with the egress of SyntheticPatientGenerator.generate fused on the device (utils/generate.py:127-144): returns
        {'mutations': uint8 [n, mutation_dim] = samples[:, :mutation_dim] > 0.5, 'expression': fp32 [n, expression_dim],
        'pathways': fp32 [n, pathway_dim], 'conditions'} as device tensors (plus 'mutation_bits', one bit per gene, LSB first, when
        pack_bits=True) instead of the dense [n, D] matrix. The calls are bit-identical to thresholding sample()'s output."""
        return self.sample(conditions, num_samples, _components="bits" if pack_bits else "bytes", **kw)

    @torch.no_grad()
    @_on_model_device
    def predict_noise(self, x_t, t, conditions):
        """eps_theta(x_t, t, c) for integer timesteps t [n] — the denoiser alone (DiffusionUNet.forward, :210-256)."""
        dev = self._device()
        x_t = self._as_f32(x_t, dev)
        conditions = self._as_f32(conditions, dev)
        n = x_t.shape[0]
        lib = self._ensure_ctx(n)
        s = _lib.stream_handle()
        t32 = t.to(device=dev, dtype=torch.int32).contiguous()
        eps = torch.empty_like(x_t)
        _lib.check(lib.osteo_ddpm_set_conditions(self._ctx, conditions.data_ptr(), n, s))
        _lib.check(lib.osteo_ddpm_denoise(self._ctx, x_t.data_ptr(), t32.data_ptr(), n, eps.data_ptr(), s))
        return eps
