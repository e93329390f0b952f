"""GPU: the rest of BiologicalValidator (utils/validation.py:27-123 co-occurrence, :225-271 statistical tests, :300-387 validate_all)
through the CUDA path against the reference's own outputs (tests/golden/validate_all.npz)."""
import numpy as np
import pytest
import torch

from oracle import validator_inputs as VI
from osteosarcoma_diffusionmodel_b200.validation import BiologicalValidator

pytestmark = pytest.mark.gpu


def test_cooccurrence_matches_reference(golden_dir):
    g = np.load(golden_dir / "validate_all.npz")
    real, syn = VI.mutation_frames()
    np.random.seed(123)
    res = BiologicalValidator(VI.CONFIG).validate_mutation_cooccurrence(real, syn)
    assert len(res) == 4
    for k, v in res.items():
        assert abs(v - float(g[f"cooc_{k}"])) < 1e-10, k          # counts are exact integers: identical chi-square scores


def test_cooccurrence_counts_are_exact_on_many_rows():
    rs = np.random.RandomState(3)
    m = (rs.random_sample((300_001, 50)) < 0.2).astype(np.float32)
    got = BiologicalValidator(VI.CONFIG)._cooccurrence_counts(torch.from_numpy(m).cuda())
    assert np.array_equal(got, m.astype(np.float64).T @ m.astype(np.float64))


def test_binary_check():
    import pandas as pd
    df = pd.DataFrame(np.array([[0, 1, 2], [1, 0, 1.0]]), columns=list("abc"))
    with pytest.raises(ValueError, match="binary"):
        BiologicalValidator(VI.CONFIG).validate_mutation_cooccurrence(df, df)


def test_statistical_tests_match_reference(golden_dir):
    g = np.load(golden_dir / "validate_all.npz")
    real, syn = VI.stat_matrices()
    res = BiologicalValidator(VI.CONFIG).statistical_tests(real, syn)
    assert abs(res["ks_test_mean_pvalue"] - float(g["stat_ks_test_mean_pvalue"])) < 1e-6
    assert res["ks_test_fraction_significant"] == float(g["stat_ks_test_fraction_significant"])
    assert abs(res["mmd"] - float(g["stat_mmd"])) < 1e-4 * float(g["stat_mmd"])
    # reference = sklearn's randomized PCA; ours = exact leading axes of an fp32x3 Gram
    assert abs(res["wasserstein_distance_mean"] - float(g["stat_wasserstein_distance_mean"])) < 2e-3 * float(g["stat_wasserstein_distance_mean"])


def test_pca_feature_major_branch_matches_row_major():
    """More patients than features: the scatter matrix is X^T X (wgrad-shaped, rows split over CTAs); same scores up to sign."""
    real, syn = VI.stat_matrices(n_real=900, n_syn=400, d=96, seed=5)
    val = BiologicalValidator(VI.CONFIG)
    r, s = torch.from_numpy(real).float().cuda(), torch.from_numpy(syn).float().cuda()
    pr, ps = val._pca_scores(r, s, 10)
    mean = real.mean(0)
    u, sv, vt = np.linalg.svd(real - mean, full_matrices=False)
    ref_r, ref_s = (real - mean) @ vt[:10].T, (syn - mean) @ vt[:10].T
    for got, ref in ((pr, ref_r), (ps, ref_s)):
        got = got.cpu().numpy().astype(np.float64)
        sign = np.sign((got * ref).sum(0))
        assert np.abs(got * sign - ref).max() < 1e-3 * np.abs(ref).max()
    w = val._wasserstein(pr, ps)
    from scipy import stats
    ref_w = [stats.wasserstein_distance(ref_r[:, i], ref_s[:, i]) for i in range(10)]
    assert np.allclose(w, ref_w, rtol=1e-3)


def test_ks_asymptotic_branch(golden_dir):
    g = np.load(golden_dir / "validate_all.npz")
    rs = np.random.RandomState(31)
    a, b = rs.standard_normal((12000, 3)), rs.standard_normal((10500, 3)) * 1.02 + 0.01
    val = BiologicalValidator(VI.CONFIG)
    d = val._ks_statistics(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda())      # float64 in, float64 sort
    assert np.allclose(d, g["ks_big_stat"], atol=1e-12)


def test_validate_all_matches_reference(golden_dir):
    g = np.load(golden_dir / "validate_all.npz")
    np.random.seed(77)
    res = BiologicalValidator(VI.CONFIG).validate_all(*VI.validate_all_frames())
    assert list(res.keys()) == list(g["all_keys"])
    for k, v in res.items():
        ref = float(g[f"all_{k}"])
        tol = 2e-3 * abs(ref) if k == "wasserstein_distance_mean" else (1e-4 * abs(ref) if k == "mmd" else 1e-6)
        assert abs(v - ref) < tol, (k, v, ref)
