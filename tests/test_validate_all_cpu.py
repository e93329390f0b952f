"""CPU: the HOST logic of the validate_all family (chi-square from counts, KS p-values, CDF integrals, PCA algebra, result keys and
the overall score) against the reference's own outputs (tests/golden/validate_all.npz, written by oracle/gen_golden.py), with the
device contractions replaced by torch CPU stand-ins. The CUDA contractions themselves are checked in tests/test_validate_all_gpu.py."""
import numpy as np
import pytest
import torch

from oracle import validator_inputs as VI
from oracle import validators_oracle as V
from osteosarcoma_diffusionmodel_b200.validation import BiologicalValidator


@pytest.fixture()
def cpu_validator(monkeypatch):
    val = BiologicalValidator(VI.CONFIG, device="cpu")
    monkeypatch.setattr(BiologicalValidator, "_require_cuda", lambda self: None)
    monkeypatch.setattr(BiologicalValidator, "_cooccurrence_counts", lambda self, m: (m.double().t() @ m.double()).numpy())
    monkeypatch.setattr(BiologicalValidator, "_gram_rows", staticmethod(lambda x: (x.double() @ x.double().t()).float()))
    monkeypatch.setattr(BiologicalValidator, "_gram_cols", staticmethod(lambda x: (x.double().t() @ x.double()).float()))
    monkeypatch.setattr(BiologicalValidator, "_project", staticmethod(lambda x, a: (x.double() @ a.double().t()).float()))
    monkeypatch.setattr(BiologicalValidator, "compute_mmd", lambda self, X, Y, kernel="rbf", gamma=None: V.compute_mmd(np.asarray(X), np.asarray(Y), gamma=gamma))

    def corr_rule(self, mutations, expression, pathway_scores):          # Series.corr stand-in for the moment kernel
        viol = tot = 0
        for rule in self.required_correlations:
            g, p = rule["mutation"], rule["pathway"]
            if g not in mutations.columns or p not in pathway_scores.columns:
                continue
            c = np.corrcoef(mutations[g].values, pathway_scores[p].values)[0, 1]
            viol += int((rule["direction"] == "positive" and c < 0) or (rule["direction"] == "negative" and c > 0))
            tot += 1
        return {"mutation_expression_violation_rate": viol / tot} if tot else {}

    monkeypatch.setattr(BiologicalValidator, "validate_mutation_expression_correlation", corr_rule)
    return val


def test_cooccurrence_matches_reference(golden_dir, cpu_validator):
    g = np.load(golden_dir / "validate_all.npz")
    real, syn = VI.mutation_frames()
    np.random.seed(123)
    res = cpu_validator.validate_mutation_cooccurrence(real, syn)
    assert set(res) == {"mutation_frequency_correlation", "driver_gene_frequency_diff", "mutual_exclusivity_violation_rate", "cooccurrence_pattern_correlation"}
    for k, v in res.items():
        assert abs(v - float(g[f"cooc_{k}"])) < 1e-10, k


def test_statistical_tests_match_reference(golden_dir, cpu_validator):
    g = np.load(golden_dir / "validate_all.npz")
    real, syn = VI.stat_matrices()
    res = cpu_validator.statistical_tests(real, syn)
    assert abs(res["ks_test_mean_pvalue"] - float(g["stat_ks_test_mean_pvalue"])) < 1e-6        # inputs pass through fp32 on the device path
    assert res["ks_test_fraction_significant"] == float(g["stat_ks_test_fraction_significant"])
    assert abs(res["mmd"] - float(g["stat_mmd"])) < 1e-6
    # the reference's PCA is sklearn's RANDOMIZED solver (utils/validation.py:256-258): exact axes agree to its own accuracy
    assert abs(res["wasserstein_distance_mean"] - float(g["stat_wasserstein_distance_mean"])) < 2e-3 * float(g["stat_wasserstein_distance_mean"])


def test_ks_asymptotic_branch(golden_dir, cpu_validator):
    g = np.load(golden_dir / "validate_all.npz")
    rs = np.random.RandomState(31)
    a, b = rs.standard_normal((12000, 3)), rs.standard_normal((10500, 3)) * 1.02 + 0.01
    d = cpu_validator._ks_statistics(torch.from_numpy(a), torch.from_numpy(b))
    assert np.allclose(d, g["ks_big_stat"], atol=1e-12)
    p = [cpu_validator._ks_pvalue(float(x), 12000, 10500) for x in d]
    assert np.allclose(p, g["ks_big_pvalue"], rtol=1e-9)


def test_validate_all_keys_and_score(golden_dir, cpu_validator):
    g = np.load(golden_dir / "validate_all.npz")
    np.random.seed(77)
    res = cpu_validator.validate_all(*VI.validate_all_frames())
    assert list(res.keys()) == list(g["all_keys"])
    for k, v in res.items():
        ref = float(g[f"all_{k}"])
        tol = 2e-3 * abs(ref) if k == "wasserstein_distance_mean" else 1e-6
        assert abs(v - ref) < tol, (k, v, ref)


def test_chi2_degenerate_tables():
    from scipy import stats
    import pandas as pd
    f = BiologicalValidator._chi2_2x2
    rs = np.random.RandomState(0)
    for _ in range(50):
        n = int(rs.randint(5, 60))
        a, b = (rs.random_sample(n) < rs.random_sample()).astype(float), (rs.random_sample(n) < rs.random_sample()).astype(float)
        ref = stats.chi2_contingency(pd.crosstab(pd.Series(a), pd.Series(b)))[0]
        assert abs(f(float((a * b).sum()), float(a.sum()), float(b.sum()), float(n)) - ref) < 1e-9
