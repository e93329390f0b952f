"""Multi-task training loss for the diffusion model (BASELINE.json configs[3]):

    total = diffusion MSE + w_pathway * pathway_coherence + w_mut_expr * mutation_expression + w_survival * survival_MSE

The reference advertises this loss (README.md:137-141) but implements it only on the cVAE wrapper, where the two biology terms are
stubs returning 0.0 (models/cvae.py:262-302) and only the survival head is real (models/cvae.py:250-255, :327-329); weights come
from config['model']['constraints'] and are combined as in models/cvae.py:334-339.  There is therefore NO reference output to match
("parity unpinned"): the differentiable forms are defined here, with forward values tied to the reference's validators:

  pathway_coherence    mean over pathways of  1 - (mean pairwise Pearson correlation of the pathway's member genes over the batch),
                       i.e. 1 - the per-pathway score of BiologicalValidator.validate_pathway_coherence (utils/validation.py:150-157);
  mutation_expression  mean over the rules of evaluation.required_correlations of  max(0, -sign * Pearson(mutation, pathway score)),
                       positive exactly when validate_mutation_expression_correlation counts a violation (utils/validation.py:206-214);
  survival             the reference's head, Linear(latent_dim, 128) -> ReLU -> Dropout(0.2) -> Linear(128, 1), MSE against the
                       (normalised) survival time.  The cVAE feeds it the latent mean; a diffusion model has none, so the head reads
                       the predicted clean sample's mutation and pathway-score blocks, zero padded / truncated to latent_dim.

All three are functions of the PREDICTED CLEAN SAMPLE x0hat = (x_t - sqrt(1 - ab_t) eps_hat) / sqrt(ab_t) (models/diffusion.py:401-403),
the diffusion analogue of the cVAE's x_recon, over the rows whose timestep still carries signal (ab_t >= min_alpha_bar: at large t
the 1 / sqrt(ab_t) factor, 6e4 at t = 999, turns x0hat into amplified noise).  The correlation losses are hand-written kernels
(validators.cuh: one-pass batched moments, a finish kernel, a row-parallel backward); their gradient reaches the denoiser through
osteo_ddpm_train_inject between the two halves of the training step.  Oracle: oracle/bio_losses_oracle.py (torch autograd, fp64).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import distributed as D
from .diffusion import BiologyAwareDiffusionModel
from .validation import _CM_STRIDE, _moments_batched


class _CorrLoss(torch.autograd.Function):
    """losses[s] of the column sets `ci` [S, 32] (int32, -1 padded) of data [n, G]; modes[s] = 0 (coherence) or +-1 (required sign)."""

    @staticmethod
    def forward(ctx, data, ci, modes, n_eff=None):
        """n_eff (optional device scalar): the number of rows that count. The other rows must be EXACTLY ZERO in `data`; the moments are then
        taken with a zero shift, to which such rows contribute nothing, and only the row count is replaced -- a row filter without
        compaction, i.e. without a device-to-host synchronisation for the size of the compacted matrix."""
        data = data.contiguous()
        lib = _lib.load()
        n, G = data.shape
        S = ci.shape[0]
        rank, ws = D.world()
        with torch.cuda.device(data.device):
            return _CorrLoss._forward_on_device(ctx, data, ci, modes, n_eff, lib, n, S, ws)

    @staticmethod
    def _forward_on_device(ctx, data, ci, modes, n_eff, lib, n, S, ws):
        # a shift near the column means conditions the fp64 moments; ranks must agree on it before their moments are summed
        if ws == 1 and n_eff is None:
            shift = data[0, ci.clamp(min=0).long()].contiguous()
        else:
            shift = torch.zeros((S, 32), device=data.device, dtype=torch.float32)
        mom = _moments_batched(data, ci, shift, (0, n))
        if n_eff is not None:
            mom[:, 0] = n_eff.to(torch.float64)
        mom = D.all_reduce_sum_(mom)
        losses = torch.empty(S, device=data.device, dtype=torch.float32)
        coef = torch.empty(S * 32 * 4, device=data.device, dtype=torch.float32)
        _lib.check(lib.osteo_corr_loss_finish(mom.data_ptr(), ci.data_ptr(), shift.data_ptr(), modes.data_ptr(), S, losses.data_ptr(), coef.data_ptr(),
                                              _lib.stream_handle()))
        ctx.save_for_backward(data, ci, coef)
        ctx.scale = float(ws)       # data-parallel averaging divides by the world size; the loss is already the global batch's
        return losses

    @staticmethod
    def backward(ctx, grad_out):
        data, ci, coef = ctx.saved_tensors
        up = (grad_out.to(torch.float32) * ctx.scale).contiguous()
        grad = torch.zeros_like(data)
        with torch.cuda.device(data.device):
            _lib.check(_lib.load().osteo_corr_loss_backward(data.data_ptr(), data.shape[0], data.shape[1], ci.data_ptr(), ci.shape[0], coef.data_ptr(), up.data_ptr(),
                                                            grad.data_ptr(), _lib.stream_handle()))
        return grad, None, None, None


def pack_column_sets(column_sets: Sequence[Sequence[int]], modes: Sequence[int], device) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """Device-side index tensors of correlation_losses, in groups of at most 32 sets: [(ci int32 [S, 32] (-1 padded), modes int32 [S])].
    Build them once for a fixed set list (BiologyConstrainedDiffusion does): every rebuild is a host-to-device copy per step."""
    if len(column_sets) != len(modes):
        raise ValueError("one mode per column set")
    packs = []
    for b0 in range(0, len(column_sets), 32):
        sets = column_sets[b0:b0 + 32]
        ci = torch.full((len(sets), 32), -1, dtype=torch.int32)
        for i, cols in enumerate(sets):
            if not 2 <= len(cols) <= 32:
                raise ValueError("a column set needs 2..32 columns")
            if modes[b0 + i] != 0 and len(cols) != 2:
                raise ValueError("a required-sign rule is a pair of columns")
            ci[i, :len(cols)] = torch.tensor(list(cols), dtype=torch.int32)
        packs.append((ci.to(device), torch.tensor(list(modes[b0:b0 + 32]), dtype=torch.int32, device=device)))
    return packs


def correlation_losses(data: torch.Tensor, column_sets: Sequence[Sequence[int]] = (), modes: Sequence[int] = (), packed=None,
                       n_eff: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Differentiable per-set losses (fp32 [n_sets]) of data [n, G] (fp32, CUDA): mode 0 = 1 - mean pairwise Pearson correlation of the
    set's columns; mode +1 / -1 = max(0, -mode * Pearson(col 0, col 1)). `packed` = a cached pack_column_sets(...) result; `n_eff` = device
    scalar, the number of rows that count when the others have been zeroed (see _CorrLoss.forward)."""
    if data.device.type != "cuda":
        raise RuntimeError("correlation_losses computes only on a CUDA device; there is no CPU fallback")
    if data.shape[1] * 4 > 96 * 1024:
        raise ValueError("gather the columns first: a row must fit the kernel's 96 KB staging buffer")
    if packed is None:
        packed = pack_column_sets(column_sets, modes, data.device)
    data = data.to(torch.float32)
    return torch.cat([_CorrLoss.apply(data, ci, md, n_eff) for ci, md in packed])


class _AuxEvaluator:
    """Evaluates the auxiliary terms and their gradient on x0hat between the two halves of a training step."""

    def __init__(self, owner: "BiologyConstrainedDiffusion", survival_time: Optional[torch.Tensor]):
        self.o = owner
        self.survival_time = survival_time
        self.columns = owner._aux_columns
        self.total = None
        self.parts: Dict[str, torch.Tensor] = {}
        self.head_grads: Optional[Tuple[torch.Tensor, ...]] = None

    def gradient(self, x0hat: torch.Tensor, t: torch.Tensor) -> Optional[torch.Tensor]:
        o = self.o
        dev = x0hat.device
        zero = torch.zeros((), device=dev)
        self.parts = {"pathway_coherence": zero, "mutation_expression": zero, "survival": zero}
        self.total = zero
        # rows whose timestep still carries signal; filtered by zeroing (no compaction: nothing here waits for the device)
        kf = (o.diffusion.alphas_cumprod[t.long()] >= o.min_alpha_bar).to(torch.float32)
        n_eff = kf.sum()
        head = list(o.survival_predictor.parameters())
        with torch.enable_grad():
            xh = x0hat.detach().requires_grad_(True)
            sub = xh * kf[:, None]
            total = zero
            if o._sets:
                losses = correlation_losses(sub, packed=o._packed_sets(dev), n_eff=n_eff)
                P = o._n_pathways
                if P and o.pathway_coherence_weight:
                    self.parts["pathway_coherence"] = losses[:P].mean()
                    total = total + o.pathway_coherence_weight * self.parts["pathway_coherence"]
                if len(o._sets) > P and o.mutation_expr_weight:
                    self.parts["mutation_expression"] = losses[P:].mean()
                    total = total + o.mutation_expr_weight * self.parts["mutation_expression"]
            if self.survival_time is not None and o.survival_weight:
                u = sub[:, o._head_pos]
                if u.shape[1] < o.latent_dim:
                    u = F.pad(u, (0, o.latent_dim - u.shape[1]))
                pred = o.survival_predictor(u).squeeze(-1)
                err = (pred - self.survival_time.to(dev, torch.float32)) ** 2
                self.parts["survival"] = (err * kf).sum() / n_eff.clamp(min=1.0)          # F.mse_loss over the kept rows
                total = total + o.survival_weight * self.parts["survival"]
            if not total.requires_grad:
                return None
            grads = torch.autograd.grad(total, [xh] + head, allow_unused=True)
        self.total = total.detach()
        self.parts = {k: v.detach() for k, v in self.parts.items()}
        self.head_grads = tuple(g if g is not None else torch.zeros_like(p) for g, p in zip(grads[1:], head))
        return grads[0]


class _MultiTaskStep(torch.autograd.Function):
    """total loss with the whole forward + backward done inside; autograd only scales the precomputed gradients."""

    @staticmethod
    def forward(ctx, owner, x0, cond, survival_time, *params):
        model = owner.diffusion
        aux = _AuxEvaluator(owner, survival_time)
        inject, model._inject = model._inject, None      # parity tests inject t / noise / dropout masks like BiologyAwareDiffusionModel
        loss, grads = model._run_train_step(x0, cond, inject, want_grads=True, aux=aux)
        head = list(owner.survival_predictor.parameters())
        ctx.model, ctx.epoch = model, model._grad_epoch
        ctx.grads = list(grads) + list(aux.head_grads if aux.head_grads is not None else [torch.zeros_like(p) for p in head])
        owner.last_losses = dict(aux.parts, diffusion=loss.detach())
        return loss + aux.total

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.model._grad_epoch != ctx.epoch:
            raise RuntimeError("backward() of a loss whose gradients were overwritten by a later forward(): call backward() before the next "
                               "training forward of this model")
        return (None, None, None, None) + tuple(torch._foreach_mul(ctx.grads, grad_out))


class BiologyConstrainedDiffusion(nn.Module):
    """Diffusion counterpart of BiologyConstrainedVAE (models/cvae.py:222-346): the DDPM plus the three auxiliary terms.

    pathway_members       column index lists INTO THE EXPRESSION BLOCK (one list per pathway, 3..32 genes: the gene-set columns
                          of utils/pathway_features.py present in the expression matrix);
    correlation_rules     (mutation column in the mutation block, pathway column in the pathway-score block, +1 | -1) triples, the
                          tensor form of config.evaluation.required_correlations (config/config.yaml:110-116).
    The wrapped model is `self.diffusion` (not `vae`: utils/train.py:233 dispatches on that attribute name), so the reference's
    Trainer drives this class through forward(x, conditions, return_loss=True)."""

    def __init__(self, mutation_dim: int, expression_dim: int, pathway_dim: int, condition_dim: int, config: dict,
                 pathway_members: Optional[Sequence[Sequence[int]]] = None, correlation_rules: Optional[Sequence[Tuple[int, int, int]]] = None):
        super().__init__()
        self.diffusion = BiologyAwareDiffusionModel(mutation_dim, expression_dim, pathway_dim, condition_dim, config)
        self.mutation_dim, self.expression_dim, self.pathway_dim, self.condition_dim = mutation_dim, expression_dim, pathway_dim, condition_dim
        self.data_dim = self.diffusion.data_dim
        self.num_steps = self.diffusion.num_steps
        mcfg = config["model"]
        self.latent_dim = int(mcfg["latent_dim"])
        self.survival_predictor = nn.Sequential(nn.Linear(self.latent_dim, 128), nn.ReLU(), nn.Dropout(0.2), nn.Linear(128, 1))    # models/cvae.py:250-255
        cons = mcfg.get("constraints", {})
        self.pathway_coherence_weight = float(cons.get("pathway_coherence_weight", 1.0))        # models/cvae.py:258-260
        self.mutation_expr_weight = float(cons.get("mutation_expression_weight", 0.5))
        self.survival_weight = float(cons.get("survival_prediction_weight", 0.3))
        b200 = mcfg.get("b200", {}) if isinstance(mcfg.get("b200", {}), dict) else {}
        self.min_alpha_bar = float(b200.get("aux_min_alpha_bar", 0.5))
        self.last_losses: Dict[str, torch.Tensor] = {}

        members = [list(m) for m in (pathway_members or [])]
        rules = [tuple(r) for r in (correlation_rules or [])]
        for m in members:
            if not 3 <= len(m) <= 32 or min(m) < 0 or max(m) >= expression_dim:
                raise ValueError("a pathway needs 3..32 member columns inside the expression block")
        for mc, pc, sign in rules:
            if not (0 <= mc < mutation_dim and 0 <= pc < pathway_dim and sign in (1, -1)):
                raise ValueError("a rule is (mutation column, pathway column, +1 | -1)")
        abs_sets = [[mutation_dim + g for g in m] for m in members] + [[mc, mutation_dim + expression_dim + pc] for mc, pc, _ in rules]
        head_cols = (list(range(mutation_dim)) + list(range(mutation_dim + expression_dim, self.data_dim)))[:self.latent_dim]
        cols = sorted(set(c for s in abs_sets for c in s) | set(head_cols))
        pos = {c: i for i, c in enumerate(cols)}
        self._sets: List[List[int]] = [[pos[c] for c in s] for s in abs_sets]         # positions inside the gathered x0hat matrix
        self._modes: List[int] = [0] * len(members) + [int(sign) for _, _, sign in rules]
        self._n_pathways = len(members)
        self._packs = None          # (device, pack_column_sets(...)): index tensors uploaded once
        self.register_buffer("_aux_columns", torch.tensor(cols, dtype=torch.int32), persistent=False)
        self.register_buffer("_head_pos", torch.tensor([pos[c] for c in head_cols], dtype=torch.long), persistent=False)

    def _packed_sets(self, device):
        if self._packs is None or self._packs[0] != device:
            self._packs = (device, pack_column_sets(self._sets, self._modes, device))
        return self._packs[1]

    def forward(self, x, conditions, return_loss: bool = True, survival_time: Optional[torch.Tensor] = None):
        """Total multi-task loss (models/cvae.py:304-341). Under no_grad / eval, or with return_loss=False, this is the wrapped
        model's forward (the auxiliary terms exist to shape gradients); `last_losses` holds the parts of the last training call."""
        if not return_loss or not (torch.is_grad_enabled() and self.training):
            return self.diffusion(x, conditions, return_loss=return_loss)
        m = self.diffusion
        dev = m._device()
        x = m._as_f32(x, dev)
        conditions = m._as_f32(conditions, dev)
        params = m._param_list() + list(self.survival_predictor.parameters())
        return _MultiTaskStep.apply(self, x, conditions, survival_time, *params)

    @torch.no_grad()
    def sample(self, conditions, num_samples: int = 1, **kw):
        return self.diffusion.sample(conditions, num_samples, **kw)      # models/cvae.py:343-346
