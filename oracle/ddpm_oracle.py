"""ORACLE — test infrastructure, not product code.

CPU restatement (torch CPU ops, fp32 unless noted) of the reference's conditional-DDPM path,
written function by function against /root/reference/models/diffusion.py.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``
may import this package; the product path (osteosarcoma_diffusionmodel_b200/) never does.

Parity pinning: the reference ships no tests or golden vectors.  This
restatement is pinned against outputs of the REFERENCE ITSELF, produced in the build container by
``oracle/gen_golden.py`` (which imports /root/reference under a ``torch_geometric`` stub) and
committed under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function below
against those fixtures (bit-exact on CPU: both sides run the same ATen kernels).

All functions take the reference's ``state_dict`` (name -> tensor) so that they are independent
of any nn.Module.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- schedule
def beta_schedule(schedule_type: str, num_steps: int) -> Tensor:
    """models/diffusion.py:312-326 (_get_beta_schedule)."""
    if schedule_type == "linear":
        return torch.linspace(1e-4, 0.02, num_steps)
    if schedule_type == "cosine":
        steps = torch.arange(num_steps + 1, dtype=torch.float32) / num_steps
        ac = torch.cos((steps + 0.008) / 1.008 * np.pi / 2) ** 2
        ac = ac / ac[0]
        betas = 1 - (ac[1:] / ac[:-1])
        return torch.clip(betas, 0.0001, 0.9999)
    raise ValueError(f"Unknown schedule: {schedule_type}")


def schedule_buffers(schedule_type: str, num_steps: int) -> Dict[str, Tensor]:
    """models/diffusion.py:299-310: the four registered buffers (fp32 cumprod)."""
    betas = beta_schedule(schedule_type, num_steps)
    alphas = 1.0 - betas
    ac = torch.cumprod(alphas, dim=0)
    return {
        "betas": betas,
        "alphas_cumprod": ac,
        "sqrt_alphas_cumprod": torch.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - ac),
    }


def reverse_coefficients(betas: Tensor, alphas_cumprod: Tensor):
    """Collapse models/diffusion.py:400-423 to x' = c_x*x - c_eps*eps + sigma*z.

    Evaluated in fp64 from the reference's fp32 buffers through its own two-term posterior mean
    (NOT the textbook 1/sqrt(alpha_t) identity, which differs at t = 1, 2 because alphas_cumprod is
    an fp32 cumprod).  sigma[0] = 0 and (c_x, c_eps)[0] encode the t == 0 branch
    x' = x0_pred.  Returns three float64 numpy arrays of length T.
    """
    b = betas.double().numpy()
    ab = alphas_cumprod.double().numpy()
    T = len(b)
    cx = np.zeros(T)
    ce = np.zeros(T)
    sg = np.zeros(T)
    for t in range(T):
        sqrt_ab = math.sqrt(ab[t])
        s1m = math.sqrt(1.0 - ab[t])
        if t == 0:
            cx[t] = 1.0 / sqrt_ab
            ce[t] = s1m / sqrt_ab
            sg[t] = 0.0
            continue
        ab_prev = ab[t - 1]
        alpha_t = 1.0 - b[t]
        k0 = math.sqrt(ab_prev) * b[t] / (1.0 - ab[t])          # multiplies x0_pred
        k1 = math.sqrt(alpha_t) * (1.0 - ab_prev) / (1.0 - ab[t])  # multiplies x_t
        cx[t] = k0 / sqrt_ab + k1
        ce[t] = k0 * s1m / sqrt_ab
        sg[t] = math.sqrt((1.0 - ab_prev) / (1.0 - ab[t]) * b[t])
    return cx, ce, sg


# --------------------------------------------------------------------------- embeddings
def time_embedding(t_norm: Tensor, dim: int) -> Tensor:
    """models/diffusion.py:124-139 (TimeEmbedding.forward); t_norm already in [0, 1)."""
    half = dim // 2
    e = np.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half, device=t_norm.device) * -e)
    e = t_norm[:, None] * e[None, :]
    return torch.cat([torch.sin(e), torch.cos(e)], dim=-1)


def time_embedding_table(num_steps: int, dim: int) -> Tensor:
    """Embedding of every integer timestep as p_sample / forward build it:
    p_sample: torch.full((B,), t / T) -> fp32 (models/diffusion.py:392);
    forward:  t.float() / T            (models/diffusion.py:367).  Both round to the same fp32
    for every t in [0, 1000) (checked in tests/test_oracle_golden.py)."""
    t_norm = torch.tensor([t / num_steps for t in range(num_steps)], dtype=torch.float32)
    return time_embedding(t_norm, dim)


def cond_embed(sd: Dict[str, Tensor], conditions: Tensor) -> Tensor:
    """models/diffusion.py:101-114 (ConditionalEmbedding)."""
    h = F.linear(conditions, sd["condition_embed.mlp.0.weight"], sd["condition_embed.mlp.0.bias"])
    h = F.silu(h)
    return F.linear(h, sd["condition_embed.mlp.2.weight"], sd["condition_embed.mlp.2.bias"])


# --------------------------------------------------------------------------- denoiser
def block_names(n_hidden: int) -> List[str]:
    """Block prefixes in execution order (models/diffusion.py:171-193, :234-251)."""
    names = [f"unet.encoder.{i}" for i in range(n_hidden - 1)]
    names.append("unet.bottleneck")
    names += [f"unet.decoder.{i}" for i in range(n_hidden - 1)]
    return names


def n_hidden_of(sd: Dict[str, Tensor]) -> int:
    return 1 + sum(1 for k in sd if k.startswith("unet.encoder.") and k.endswith(".0.weight"))


def _block(sd, prefix: str, h: Tensor, drop_mask: Optional[Tensor], p: float, training: bool) -> Tensor:
    """models/diffusion.py:198-208 (_make_block): Linear, GroupNorm(8), SiLU, Dropout, Linear, GroupNorm(8), SiLU."""
    h = F.linear(h, sd[f"{prefix}.0.weight"], sd[f"{prefix}.0.bias"])
    h = F.group_norm(h, 8, sd[f"{prefix}.1.weight"], sd[f"{prefix}.1.bias"], 1e-5)
    h = F.silu(h)
    if training and p > 0.0:
        if drop_mask is None:
            raise ValueError("training-mode oracle needs injected dropout keep-masks")
        h = h * drop_mask.to(h.dtype) / (1.0 - p)
    h = F.linear(h, sd[f"{prefix}.4.weight"], sd[f"{prefix}.4.bias"])
    h = F.group_norm(h, 8, sd[f"{prefix}.5.weight"], sd[f"{prefix}.5.bias"], 1e-5)
    return F.silu(h)


def unet_forward(sd, x: Tensor, t_norm: Tensor, c_emb: Tensor, drop_masks: Optional[Sequence[Tensor]] = None,
                 p: float = 0.0, training: bool = False) -> Tensor:
    """models/diffusion.py:210-256 (DiffusionUNet.forward)."""
    time_dim = sd["unet.time_proj.weight"].shape[1]
    t_emb = F.linear(time_embedding(t_norm, time_dim), sd["unet.time_proj.weight"], sd["unet.time_proj.bias"])
    c = F.linear(c_emb, sd["unet.cond_proj.weight"], sd["unet.cond_proj.bias"])
    h = F.linear(x, sd["unet.input_proj.weight"], sd["unet.input_proj.bias"])
    h = h + t_emb + c
    nh = n_hidden_of(sd)
    names = block_names(nh)
    masks = list(drop_masks) if drop_masks is not None else [None] * len(names)
    skips = []
    bi = 0
    for _ in range(nh - 1):
        h = _block(sd, names[bi], h, masks[bi], p, training)
        skips.append(h)
        bi += 1
    h = _block(sd, names[bi], h, masks[bi], p, training)
    bi += 1
    for _ in range(nh - 1):
        if not skips:
            break
        h = torch.cat([h, skips.pop()], dim=-1)
        h = _block(sd, names[bi], h, masks[bi], p, training)
        bi += 1
    return F.linear(h, sd["unet.output_proj.weight"], sd["unet.output_proj.bias"])


# --------------------------------------------------------------------------- diffusion
def q_sample(sd, x0: Tensor, t: Tensor, noise: Tensor) -> Tensor:
    """models/diffusion.py:328-342."""
    a = sd["sqrt_alphas_cumprod"][t].view(-1, 1)
    b = sd["sqrt_one_minus_alphas_cumprod"][t].view(-1, 1)
    return a * x0 + b * noise


def predict_eps(sd, x_t: Tensor, t: Tensor, conditions: Tensor, num_steps: int, drop_masks=None, p: float = 0.0,
                training: bool = False) -> Tensor:
    """models/diffusion.py:367-373: t integer tensor [B]."""
    t_norm = t.float() / num_steps
    return unet_forward(sd, x_t, t_norm, cond_embed(sd, conditions), drop_masks, p, training)


def forward_loss(sd, x0: Tensor, conditions: Tensor, t: Tensor, noise: Tensor, num_steps: int, drop_masks=None,
                 p: float = 0.0, training: bool = False) -> Tensor:
    """models/diffusion.py:344-378 with the random draws (randint t, randn noise, dropout masks) injected."""
    x_t = q_sample(sd, x0, t, noise)
    eps = predict_eps(sd, x_t, t, conditions, num_steps, drop_masks, p, training)
    return F.mse_loss(eps, noise)


@torch.no_grad()
def p_sample(sd, x_t: Tensor, t: int, conditions: Tensor, noise: Optional[Tensor], num_steps: int, return_eps: bool = False):
    """models/diffusion.py:382-425, literally (two-term posterior mean), noise injected."""
    B = x_t.shape[0]
    t_norm = torch.full((B,), t / num_steps)
    eps = unet_forward(sd, x_t, t_norm, cond_embed(sd, conditions))
    betas, ac = sd["betas"], sd["alphas_cumprod"]
    alpha_t = 1.0 - betas[t]
    ab_t = ac[t]
    x0_pred = (x_t - torch.sqrt(1 - ab_t) * eps) / torch.sqrt(ab_t)
    if t > 0:
        ab_prev = ac[t - 1]
        mean = (torch.sqrt(ab_prev) * betas[t] * x0_pred / (1 - ab_t) + torch.sqrt(alpha_t) * (1 - ab_prev) * x_t / (1 - ab_t))
        var = (1 - ab_prev) / (1 - ab_t) * betas[t]
        out = mean + torch.sqrt(var) * noise
    else:
        out = x0_pred
    return (out, eps) if return_eps else out


@torch.no_grad()
def sample(sd, conditions: Tensor, x_T: Tensor, noises, num_steps: int, t_stop: int = 0) -> Tensor:
    """models/diffusion.py:427-449 with x_T and the per-step noises injected.
    ``noises(t)`` returns the z of step t (t > 0)."""
    x = x_T
    for t in reversed(range(t_stop, num_steps)):
        x = p_sample(sd, x, t, conditions, noises(t) if t > 0 else None, num_steps)
    return x
