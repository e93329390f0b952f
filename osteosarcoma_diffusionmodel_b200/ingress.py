"""Training ingress on the device (SURVEY.md §8f "next" #2): the step either side of the fwd/bwd kernels.

The reference keeps the dataset on the host and, per batch, collates it with a DataLoader, copies three tensors to the device and
runs MixupAugmentation with a CPU `randperm` (utils/train.py:22-126, :204-227, :423-436).  At the per-GPU batch of BASELINE.json
configs[3] (8192 x 5142 fp32 = 168 MB) that is a pageable host-to-device copy per step, several times the 1.2 ms training step.
Here the dataset lives in HBM once and a batch is ONE kernel per tensor: gather by row index with the mixup fused in
(`osteo_mixup_rows`, bit-identical to the reference's `lam * data + (1 - lam) * data[index]`).

  GpuResidentDataset   the tensors of OsteosarcomaDataset (`data` = [mutations | expression | pathways], `conditions`,
                       `survival_days`; utils/train.py:52-68) moved to the device once; `batches()` = DataLoader(shuffle, drop_last).
                       `from_frames()` builds them from the four aligned tables `prepare_data` reads (utils/train.py:342-409): the
                       z-scoring of the pathway scores and of the survival time, the index alignment and the NaN rules included.
  MixupAugmentation    same constructor and call signature as utils/train.py:85-126 for already-gathered batches, plus
                       `gather(dataset, index)` which fuses the gather and the mix.
There is no CPU fallback.
"""
from __future__ import annotations

from typing import Dict, Iterator, Optional

import numpy as np
import torch

from . import _lib


def _rows(src: torch.Tensor, idx_a: Optional[torch.Tensor], idx_b: Optional[torch.Tensor], n: int, lam: float) -> torch.Tensor:
    """out[i] = lam * src[idx_a[i]] + (1 - lam) * src[idx_b[i]] (idx_b None: plain gather) for a 1-D or 2-D fp32 CUDA tensor."""
    if src.device.type != "cuda":
        raise RuntimeError("the device-resident ingress computes only on a CUDA device; there is no CPU fallback")
    if src.dtype != torch.float32 or not src.is_contiguous():
        raise ValueError("the dataset tensors must be contiguous fp32")
    d = 1 if src.dim() == 1 else src.shape[1]
    out = torch.empty((n,) if src.dim() == 1 else (n, d), device=src.device, dtype=torch.float32)
    for ix in (idx_a, idx_b):
        if ix is not None and (ix.dtype != torch.int64 or ix.device != src.device or not ix.is_contiguous() or ix.numel() != n):
            raise ValueError("row indices must be contiguous int64 tensors of the batch size on the dataset's device")
    # torch multiplies an fp32 tensor by a Python scalar after rounding the scalar to fp32; (1 - lam) is formed in double first
    _lib.check(_lib.load().osteo_mixup_rows(src.data_ptr(), src.shape[0], d, _lib.ptr(idx_a), _lib.ptr(idx_b), n, float(np.float32(lam)),
                                            float(np.float32(1.0 - lam)), out.data_ptr(), _lib.stream_handle()))
    return out


class GpuResidentDataset:
    """OsteosarcomaDataset's tensors (utils/train.py:52-68) resident on one device."""

    def __init__(self, data: torch.Tensor, conditions: torch.Tensor, survival_days: torch.Tensor, device="cuda"):
        if not (data.shape[0] == conditions.shape[0] == survival_days.shape[0]):
            raise ValueError("data, conditions and survival_days must have the same number of rows")
        self.data = data.to(device=device, dtype=torch.float32).contiguous()
        self.conditions = conditions.to(device=device, dtype=torch.float32).contiguous()
        self.survival_days = survival_days.to(device=device, dtype=torch.float32).contiguous()

    @classmethod
    def from_dataset(cls, dataset, device="cuda") -> "GpuResidentDataset":
        """From the reference's OsteosarcomaDataset (or a torch Subset of it, as random_split returns: utils/train.py:416-420)."""
        base, idx = dataset, None
        if hasattr(dataset, "dataset") and hasattr(dataset, "indices"):
            base, idx = dataset.dataset, torch.as_tensor(list(dataset.indices), dtype=torch.long)
        pick = (lambda t: t[idx]) if idx is not None else (lambda t: t)
        return cls(pick(base.data), pick(base.conditions), pick(base.survival_days), device=device)

    @classmethod
    def from_frames(cls, mutation_matrix, expression_matrix, pathway_scores, clinical_data, condition_features=None, normalize: bool = True,
                    device="cuda") -> "GpuResidentDataset":
        """From the four tables of `prepare_data` (utils/train.py:349-363: patients x genes / pathways DataFrames indexed by submitter id, and
        the clinical table with a `submitter_id` column), doing what prepare_data + OsteosarcomaDataset do between the CSVs and the tensors:
          * pathway scores z-scored per column over ALL their rows, `(x - mean) / (std + 1e-8)`, sample std (ddof = 1, NaN skipped) like
            pandas (utils/train.py:387); `survival_days_norm` the same way from `survival_days` (:390-392); expression stays as it is (:384);
          * condition features = those of [survival_days_norm, event_occurred, age_years, metastasis_at_diagnosis] present (:395-398),
            unless given;
          * rows = the ids common to all four tables, in the mutation table's order (utils/train.py:38-43); NaN conditions -> 0, NaN
            survival -> 0 (:60-66); `data` = [mutations | expression | pathways] fp32 (:46-56).
        The statistics are taken in float64 on the device (the tables can be millions of rows) and the result is rounded to fp32 once, as
        pandas' float64 arithmetic followed by `.astype(np.float32)` does. `normalize=False` skips the two z-scorings (tables that
        prepare_data has already normalised). `.config_dims()` gives the four sizes prepare_data stores in config['model'] (:439-442)."""
        import pandas as pd

        dev = torch.device(device)
        clinical = clinical_data.set_index("submitter_id")
        common = mutation_matrix.index.intersection(expression_matrix.index).intersection(pathway_scores.index).intersection(clinical.index)

        def zscore(t: torch.Tensor) -> torch.Tensor:          # float64 [n, k] on the device, NaN-skipping, ddof = 1
            ok = ~torch.isnan(t)
            cnt = ok.sum(0).to(torch.float64)
            x = torch.where(ok, t, torch.zeros_like(t))
            mean = x.sum(0) / cnt
            var = (torch.where(ok, t - mean, torch.zeros_like(t)) ** 2).sum(0) / (cnt - 1.0)
            return (t - mean) / (var.sqrt() + 1e-8)

        def dev64(frame_or_series) -> torch.Tensor:
            a = np.array(frame_or_series.to_numpy(dtype=np.float64), copy=True, order="C")          # pandas may hand out a read-only view
            return torch.from_numpy(a).to(dev)

        paths = dev64(pathway_scores)
        if normalize:
            paths = zscore(paths)
        pos = torch.as_tensor(pathway_scores.index.get_indexer(common), device=dev)
        paths = paths[pos]
        clin = clinical.copy()
        if normalize and "survival_days" in clin.columns:
            sd = zscore(dev64(clinical_data["survival_days"]).reshape(-1, 1)).reshape(-1)
            clin["survival_days_norm"] = sd.cpu().numpy()
        if condition_features is None:
            condition_features = [f for f in ("survival_days_norm", "event_occurred", "age_years", "metastasis_at_diagnosis") if f in clin.columns]
        clin = clin.loc[common]
        cond = torch.nan_to_num(dev64(clin[list(condition_features)]).to(torch.float32), nan=0.0)
        surv = dev64(clin["survival_days"].fillna(0)).to(torch.float32)
        mut = torch.from_numpy(np.array(mutation_matrix.loc[common].to_numpy(dtype=np.float32), copy=True, order="C")).to(dev)
        expr = torch.from_numpy(np.array(expression_matrix.loc[common].to_numpy(dtype=np.float32), copy=True, order="C")).to(dev)
        out = cls(torch.cat([mut, expr, paths.to(torch.float32)], dim=1), cond, surv, device=dev)
        out.condition_features = list(condition_features)
        out._dims = {"n_genes_mutation": int(mut.shape[1]), "n_genes_expression": int(expr.shape[1]), "n_pathways": int(paths.shape[1]),
                     "n_conditions": len(condition_features)}
        out.index = pd.Index(common)
        return out

    def config_dims(self) -> Dict[str, int]:
        """The sizes prepare_data writes into config['model'] (utils/train.py:439-442); only for datasets built by from_frames()."""
        return dict(self._dims)

    def __len__(self) -> int:
        return self.data.shape[0]

    def gather(self, index: torch.Tensor, mix_index: Optional[torch.Tensor] = None, lam: float = 1.0) -> Dict[str, torch.Tensor]:
        """The batch dict of OsteosarcomaDataset.__getitem__ + default collate (utils/train.py:77-82) for dataset rows `index`,
        optionally mixed with rows `mix_index` (MixupAugmentation, utils/train.py:117-120)."""
        n = index.numel()
        return {"data": _rows(self.data, index, mix_index, n, lam), "conditions": _rows(self.conditions, index, mix_index, n, lam),
                "survival": _rows(self.survival_days, index, mix_index, n, lam)}

    def batches(self, batch_size: int, shuffle: bool = True, drop_last: bool = True, generator: Optional[torch.Generator] = None,
                mixup: Optional["MixupAugmentation"] = None) -> Iterator[Dict[str, torch.Tensor]]:
        """DataLoader(batch_size, shuffle, drop_last, num_workers=0) over the resident tensors (utils/train.py:423-436); with `mixup`
        every batch is gathered and mixed in the same kernel."""
        n = len(self)
        order = torch.randperm(n, device=self.data.device, generator=generator) if shuffle else torch.arange(n, device=self.data.device)
        stop = n - n % batch_size if drop_last else n
        for b0 in range(0, stop, batch_size):
            index = order[b0:min(b0 + batch_size, n)].contiguous()
            yield mixup.gather(self, index) if mixup is not None else self.gather(index)


class MixupAugmentation:
    """Mixup data augmentation (utils/train.py:85-126), on the device."""

    def __init__(self, alpha: float = 0.2):
        self.alpha = alpha

    def _lam(self) -> float:
        return float(np.random.beta(self.alpha, self.alpha)) if self.alpha > 0 else 1.0      # utils/train.py:107-110

    def __call__(self, batch: Dict[str, torch.Tensor], lam: Optional[float] = None, index: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """Mix an already-gathered batch {'data', 'conditions', 'survival'} (same contract as the reference; `lam` / `index` can be
        injected for parity runs, otherwise Beta(alpha, alpha) and a device-side randperm)."""
        data = batch["data"]
        n = data.shape[0]
        lam = self._lam() if lam is None else float(lam)
        if index is None:
            index = torch.randperm(n, device=data.device)                                       # utils/train.py:115 (CPU there)
        index = index.to(device=data.device, dtype=torch.int64).contiguous()
        f32 = lambda t: t.to(torch.float32).contiguous()                                        # noqa: E731
        return {k: _rows(f32(batch[k]), None, index, n, lam) for k in ("data", "conditions", "survival")}

    def gather(self, dataset: GpuResidentDataset, index: torch.Tensor, lam: Optional[float] = None, perm: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """Gather dataset rows `index` and mix them with a permutation of themselves in one pass over the data."""
        n = index.numel()
        lam = self._lam() if lam is None else float(lam)
        if perm is None:
            perm = torch.randperm(n, device=index.device)
        mix_index = index[perm.to(index.device)].contiguous()
        return dataset.gather(index.contiguous(), mix_index, lam)
