"""Deterministic synthetic inputs (weights, cohorts, noise) of the preprocessor's output shape: used by bench.py, the
golden generator and the tests. Pure data generation - no model arithmetic lives here.

Everything is drawn from numpy's frozen legacy ``RandomState`` (MT19937 streams are guaranteed
stable across numpy versions), so the fixtures under tests/golden/ only need to store OUTPUTS:
weights, cohorts and injected noise are regenerated identically on any machine.
Shapes follow the preprocessor's output: mutation [N, Dm] in {0,1},
expression [N, De] fp32 ~ N(0,1), pathway scores [N, Dp] z-scored.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

CONFIG_YAML_DIMS = dict(mutation_dim=62, expression_dim=5054, pathway_dim=26, condition_dim=3)   # config/config.yaml:27-30
SMOKE_DIMS = dict(mutation_dim=100, expression_dim=200, pathway_dim=50, condition_dim=5)         # models/diffusion.py:466-472
SCENARIO_CONDITIONS = [[2.4, 0.0, 0.0], [-1.0, 1.0, 1.0], [0.0, 0.0, 0.0]]   # utils/generate.py:61-82 on config.yaml:121-141


def model_config(hidden_dims: Sequence[int] = (256, 512, 256), num_steps: int = 1000, schedule: str = "cosine",
                 dropout: float = 0.2, latent_dim: int = 128) -> dict:
    """The five config keys the model reads (models/diffusion.py:291-300)."""
    return {"model": {"latent_dim": latent_dim, "hidden_dims": list(hidden_dims), "gnn": {"dropout": dropout},
                      "diffusion": {"num_steps": num_steps, "beta_schedule": schedule}}}


def param_shapes(data_dim: int, cond_dim: int, hidden_dims: Sequence[int], latent_dim: int = 128, cond_embed: int = 64) -> List[tuple]:
    """(name, shape) of the reference's parameters in state_dict order (SURVEY.md §8a layer table)."""
    h = list(hidden_dims)
    out = [("condition_embed.mlp.0.weight", (cond_embed, cond_dim)), ("condition_embed.mlp.0.bias", (cond_embed,)),
           ("condition_embed.mlp.2.weight", (cond_embed, cond_embed)), ("condition_embed.mlp.2.bias", (cond_embed,)),
           ("unet.input_proj.weight", (h[0], data_dim)), ("unet.input_proj.bias", (h[0],)),
           ("unet.cond_proj.weight", (h[0], latent_dim // 2)), ("unet.cond_proj.bias", (h[0],)),
           ("unet.time_proj.weight", (h[0], latent_dim)), ("unet.time_proj.bias", (h[0],))]

    def block(prefix, cin, cout):
        return [(f"{prefix}.0.weight", (cout, cin)), (f"{prefix}.0.bias", (cout,)), (f"{prefix}.1.weight", (cout,)), (f"{prefix}.1.bias", (cout,)),
                (f"{prefix}.4.weight", (cout, cout)), (f"{prefix}.4.bias", (cout,)), (f"{prefix}.5.weight", (cout,)), (f"{prefix}.5.bias", (cout,))]

    cin = h[0]
    for i, c in enumerate(h[1:]):
        out += block(f"unet.encoder.{i}", cin, c)
        cin = c
    out += block("unet.bottleneck", cin, cin)
    cur = h[-1]
    j = 0
    for i in range(len(h) - 2, -1, -1):
        out += block(f"unet.decoder.{j}", cur + h[i + 1], h[i])
        cur = h[i]
        j += 1
    out += [("unet.output_proj.weight", (data_dim, cur)), ("unet.output_proj.bias", (data_dim,))]
    return out


def make_params(data_dim: int, cond_dim: int, hidden_dims: Sequence[int] = (256, 512, 256), seed: int = 0, latent_dim: int = 128) -> Dict[str, torch.Tensor]:
    """Deterministic 'trained-looking' parameters: Linear weights ~ N(0, 1/fan_in), biases ~ 0.1 N(0,1),
    GroupNorm gamma ~ 1 + 0.1 N(0,1), beta ~ 0.1 N(0,1) (non-trivial affine so parity tests exercise it)."""
    rs = np.random.RandomState(seed)
    sd = {}
    for name, shape in param_shapes(data_dim, cond_dim, hidden_dims, latent_dim):
        leaf = name.split(".")
        is_gn = leaf[-2] in ("1", "5") and leaf[0] == "unet" and len(shape) == 1 and leaf[1] in ("encoder", "bottleneck", "decoder")
        if len(shape) == 2:
            v = rs.standard_normal(shape) / np.sqrt(shape[1])
        elif is_gn and leaf[-1] == "weight":
            v = 1.0 + 0.1 * rs.standard_normal(shape)
        else:
            v = 0.1 * rs.standard_normal(shape)
        sd[name] = torch.from_numpy(v.astype(np.float32))
    return sd


def make_cohort(n: int, mutation_dim: int, expression_dim: int, pathway_dim: int, condition_dim: int, seed: int = 0):
    """x0 = [mutations | expression | pathways] (utils/train.py:56) and condition vectors."""
    rs = np.random.RandomState(seed + 1000)
    mut = (rs.random_sample((n, mutation_dim)) < 0.1).astype(np.float32)
    expr = rs.standard_normal((n, expression_dim)).astype(np.float32)
    path = rs.standard_normal((n, pathway_dim)).astype(np.float32)
    x0 = np.concatenate([mut, expr, path], axis=1)
    cond = rs.standard_normal((n, condition_dim)).astype(np.float32)
    if condition_dim >= 2:
        cond[:, 1] = (rs.random_sample(n) < 0.4).astype(np.float32)   # event_occurred
    return torch.from_numpy(x0), torch.from_numpy(cond)


def scenario_conditions(n: int, condition_dim: int = 3) -> torch.Tensor:
    """The three config.yaml scenarios in equal thirds (main.py:227-228)."""
    rows = [SCENARIO_CONDITIONS[(3 * i) // max(n, 1)][:condition_dim] for i in range(n)]
    return torch.tensor(rows, dtype=torch.float32)


def noise_stream(seed: int):
    """Stateless per-step noise: noise_stream(seed)(tag, shape) -> fp32 tensor, independent of call order."""
    def draw(tag: int, shape) -> torch.Tensor:
        rs = np.random.RandomState((seed * 1_000_003 + tag) % (2 ** 31 - 1))
        return torch.from_numpy(rs.standard_normal(shape).astype(np.float32))
    return draw


def dropout_masks(seed: int, n: int, widths: Sequence[int], p: float) -> List[torch.Tensor]:
    """Keep-masks (1 = keep) for the Dropout of each block, in execution order."""
    rs = np.random.RandomState(seed + 77)
    return [torch.from_numpy((rs.random_sample((n, w)) >= p).astype(np.uint8)) for w in widths]


def block_widths(hidden_dims: Sequence[int]) -> List[int]:
    h = list(hidden_dims)
    return h[1:] + [h[-1]] + [h[i] for i in range(len(h) - 2, -1, -1)]
