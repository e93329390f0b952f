"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): torch-CPU restatement of the reference's batch collation and MixupAugmentation
(utils/train.py:77-82, :85-126) with the random draws (lam, index) injected."""
from __future__ import annotations

import torch


def mixup(batch: dict, lam: float, index: torch.Tensor) -> dict:
    """utils/train.py:117-126, verbatim arithmetic: Python-scalar lam times fp32 tensors."""
    data, conditions, survival = batch["data"], batch["conditions"], batch["survival"]
    return {"data": lam * data + (1 - lam) * data[index], "conditions": lam * conditions + (1 - lam) * conditions[index],
            "survival": lam * survival + (1 - lam) * survival[index]}


def collate(data: torch.Tensor, conditions: torch.Tensor, survival_days: torch.Tensor, index: torch.Tensor) -> dict:
    """OsteosarcomaDataset.__getitem__ + default_collate for the rows `index` (utils/train.py:77-82)."""
    return {"data": data[index], "conditions": conditions[index], "survival": survival_days[index]}
