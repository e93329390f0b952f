"""`GpuResidentDataset.from_frames` against the reference's own prepare_data normalisation + OsteosarcomaDataset (utils/train.py:22-75,
:342-409), on the CPU (the table preparation is plain torch: only gather / mixup need the CUDA library)."""
from pathlib import Path

import numpy as np
import pandas as pd
import pytest
import torch

from oracle import reference_import as R

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.skipif(not R.available(), reason="reference sources not staged (python -m oracle.stage_reference)")


def _tables(n=57, seed=3):
    rs = np.random.RandomState(seed)
    ids = [f"P{i:03d}" for i in range(n)]
    mut = pd.DataFrame(rs.randint(0, 2, size=(n, 7)).astype(float), index=ids, columns=[f"g{i}" for i in range(7)])
    expr = pd.DataFrame(rs.standard_normal((n, 11)) * 3 + 5, index=ids, columns=[f"e{i}" for i in range(11)])
    # the pathway table has its own row order and two extra patients; the clinical table misses one patient and has NaNs
    pids = list(rs.permutation(ids)) + ["X1", "X2"]
    path = pd.DataFrame(rs.standard_normal((n + 2, 5)) * 2 + 1, index=pids, columns=[f"p{i}" for i in range(5)])
    clin = pd.DataFrame({"submitter_id": ids[:-1], "survival_days": rs.uniform(30, 3000, n - 1), "event_occurred": rs.randint(0, 2, n - 1).astype(float),
                         "age_years": rs.uniform(5, 40, n - 1)})
    clin.loc[4, "survival_days"] = np.nan
    clin.loc[9, "age_years"] = np.nan
    return mut, expr, path, clin


def test_from_frames_matches_prepare_data_and_the_reference_dataset():
    from osteosarcoma_diffusionmodel_b200.ingress import GpuResidentDataset

    mut, expr, path, clin = _tables()
    ours = GpuResidentDataset.from_frames(mut, expr, path, clin, device="cpu")
    # the reference: prepare_data's normalisation (utils/train.py:387-398) then its dataset class
    with R.dropin_path(str(ROOT)):
        import importlib
        train = importlib.import_module("utils.train")
        p2 = (path - path.mean()) / (path.std() + 1e-8)
        c2 = clin.copy()
        c2["survival_days_norm"] = (c2["survival_days"] - c2["survival_days"].mean()) / (c2["survival_days"].std() + 1e-8)
        feats = [f for f in ["survival_days_norm", "event_occurred", "age_years", "metastasis_at_diagnosis"] if f in c2.columns]
        ref = train.OsteosarcomaDataset(mutation_matrix=mut, expression_matrix=expr, pathway_scores=p2, clinical_data=c2, condition_features=feats)
    assert ours.condition_features == feats
    assert len(ours) == len(ref) == 56
    assert ours.config_dims() == {"n_genes_mutation": 7, "n_genes_expression": 11, "n_pathways": 5, "n_conditions": 3}
    # float64 statistics summed in a different order, rounded to fp32 once: equal to the last fp32 bit or one off
    assert torch.equal(ours.data[:, :18], ref.data[:, :18])
    np.testing.assert_allclose(ours.data.numpy(), ref.data.numpy(), rtol=2e-7, atol=1e-7)
    np.testing.assert_allclose(ours.conditions.numpy(), ref.conditions.numpy(), rtol=2e-7, atol=1e-7)
    assert torch.equal(ours.survival_days, ref.survival_days)
    assert float(ours.conditions[4, 0]) == 0.0 and float(ours.conditions[9, 2]) == 0.0          # NaN -> 0 like the reference


def test_from_frames_without_normalisation_is_the_plain_dataset():
    from osteosarcoma_diffusionmodel_b200.ingress import GpuResidentDataset

    mut, expr, path, clin = _tables(n=23, seed=5)
    ours = GpuResidentDataset.from_frames(mut, expr, path, clin, condition_features=["event_occurred", "age_years"], normalize=False, device="cpu")
    with R.dropin_path(str(ROOT)):
        import importlib
        train = importlib.import_module("utils.train")
        ref = train.OsteosarcomaDataset(mutation_matrix=mut, expression_matrix=expr, pathway_scores=path, clinical_data=clin, condition_features=["event_occurred", "age_years"])
    assert torch.equal(ours.data, ref.data) and torch.equal(ours.conditions, ref.conditions) and torch.equal(ours.survival_days, ref.survival_days)
