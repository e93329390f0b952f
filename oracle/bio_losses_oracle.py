"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): torch-autograd restatement, in float64, of the differentiable biology losses of
SURVEY.md §8a row A12.  The reference has no implementation (models/cvae.py:262-302 are stubs returning 0.0): "parity unpinned".
The forward values are tied to the reference's validators instead -- 1 - (per-pathway score of validate_pathway_coherence,
utils/validation.py:150-157) and the violation test of validate_mutation_expression_correlation (utils/validation.py:206-214) --
and tests/test_multitask_gpu.py checks the CUDA kernels against this file's values and autograd gradients."""
from __future__ import annotations

from typing import Sequence

import torch


def pearson_matrix(x: torch.Tensor) -> torch.Tensor:
    """DataFrame.corr() on complete data: biased or unbiased normalisation cancels in the ratio."""
    xc = x - x.mean(0, keepdim=True)
    cov = xc.t() @ xc
    sd = torch.sqrt(torch.diagonal(cov))
    return cov / torch.outer(sd, sd)


def correlation_losses(data: torch.Tensor, column_sets: Sequence[Sequence[int]], modes: Sequence[int]) -> torch.Tensor:
    out = []
    for cols, mode in zip(column_sets, modes):
        r = pearson_matrix(data[:, list(cols)].double())
        k = len(cols)
        if mode == 0:
            iu = torch.triu_indices(k, k, offset=1)
            out.append(1.0 - r[iu[0], iu[1]].mean())          # utils/validation.py:153
        else:
            out.append(torch.relu(-float(mode) * r[0, 1]))     # violation <=> sign(corr) != expected (utils/validation.py:209-212)
    return torch.stack(out)
