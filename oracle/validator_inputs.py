"""ORACLE support — deterministic inputs of the `validate_all` family (utils/validation.py:27-123, :225-271, :300-387), shared by
oracle/gen_golden.py (which feeds them to the reference) and the tests (which feed them to the CUDA path). Test infrastructure."""
from __future__ import annotations

import numpy as np

CONFIG = {"evaluation": {"driver_genes": ["TP53", "RB1", "ATRX"], "mutually_exclusive_pairs": [["TP53", "MDM2"], ["RB1", "CDK4"]],
                         "required_correlations": [{"mutation": "TP53", "pathway": "HALLMARK_P53_PATHWAY", "direction": "negative"},
                                                   {"mutation": "MYC", "pathway": "HALLMARK_MYC_TARGETS_V1", "direction": "positive"}]}}
NAMED = ["TP53", "RB1", "MDM2", "MYC", "CDK4", "ATRX"]


def mutation_frames(n_real=240, n_syn=310, k=70, seed=21):
    """Two binary cohorts with correlated gene pairs, one constant (never mutated) column in each, and named driver genes."""
    import pandas as pd

    rs = np.random.RandomState(seed)
    genes = NAMED + [f"G{i}" for i in range(k - len(NAMED))]
    freq = 0.05 + 0.4 * rs.random_sample(k)

    def cohort(n, scale):
        latent = rs.standard_normal((n, 5))
        load = rs.standard_normal((5, k)) * 0.8
        u = latent @ load + rs.standard_normal((n, k))
        q = 1 - np.clip(freq * scale, 0.01, 0.9)
        thr = np.array([np.quantile(u[:, j], q[j]) for j in range(k)])
        m = (u > thr).astype(np.float64)
        m[:, 7] = 0.0                       # a gene nobody carries: its crosstab has one row (dof 0)
        m[:, 2] = np.where(m[:, 0] == 1, 0.0, m[:, 2])      # MDM2 mostly exclusive with TP53 ...
        m[:3, 2] = 1.0
        m[:3, 0] = 1.0                      # ... but not always
        return pd.DataFrame(m, columns=genes)

    return cohort(n_real, 1.0), cohort(n_syn, 1.15)


def stat_matrices(n_real=200, n_syn=260, d=150, seed=22):
    """Low-rank + noise cohorts with a clear spectral gap after the 10th component (so sklearn's randomized PCA agrees with the exact one)."""
    rs = np.random.RandomState(seed)
    sv = np.array([9.0, 8.0, 7.2, 6.5, 5.9, 5.2, 4.6, 4.1, 3.6, 3.1, 0.4, 0.3])
    load = rs.standard_normal((len(sv), d)) * sv[:, None] / np.sqrt(d) * 4
    real = rs.standard_normal((n_real, len(sv))) @ load + 0.25 * rs.standard_normal((n_real, d)) + 1.5
    syn = (rs.standard_normal((n_syn, len(sv))) * 1.1 + 0.15) @ load + 0.3 * rs.standard_normal((n_syn, d)) + 1.45
    return real, syn


def validate_all_frames(seed=23):
    """The six DataFrames of validate_all: mutations as above, expression [n, 120], pathway scores [n, 12] with the two rule pathways."""
    import pandas as pd

    real_mut, syn_mut = mutation_frames(180, 222, 40, seed)
    rs = np.random.RandomState(seed + 1)
    ereal, esyn = stat_matrices(180, 222, 120, seed + 2)
    pnames = ["HALLMARK_P53_PATHWAY", "HALLMARK_MYC_TARGETS_V1"] + [f"HALLMARK_X{i}" for i in range(10)]
    preal = rs.standard_normal((180, 12))
    psyn = rs.standard_normal((222, 12))
    psyn[:, 0] -= 0.7 * syn_mut["TP53"].values
    psyn[:, 1] -= 0.4 * syn_mut["MYC"].values
    egenes = [f"E{i}" for i in range(120)]
    return (real_mut, pd.DataFrame(ereal, columns=egenes), pd.DataFrame(preal, columns=pnames),
            syn_mut, pd.DataFrame(esyn, columns=egenes), pd.DataFrame(psyn, columns=pnames))
