"""GPU: MMD / pathway coherence / mutation-pathway correlation against the reference's own outputs (golden) and the numpy oracle."""
import numpy as np
import pytest
import torch

from oracle import validators_oracle as V
from osteosarcoma_diffusionmodel_b200.validation import BiologicalValidator
from tests.test_validators_oracle import _mmd_inputs, coherence_inputs, mutexpr_inputs

pytestmark = pytest.mark.gpu

CONFIG = {"evaluation": {"driver_genes": ["TP53"], "mutually_exclusive_pairs": [],
                         "required_correlations": [{"mutation": "TP53", "pathway": "HALLMARK_P53_PATHWAY", "direction": "negative"},
                                                   {"mutation": "MYC", "pathway": "HALLMARK_MYC_TARGETS_V1", "direction": "positive"}]}}


def test_mmd_matches_reference_golden(golden_dir):
    g = np.load(golden_dir / "validators.npz")
    val = BiologicalValidator(CONFIG)
    rs = np.random.RandomState(11)
    for tag in ("small", "wide"):
        n, m, d = (int(v) for v in g[f"mmd_{tag}_shape"])
        X, Y = _mmd_inputs(n, m, d, rs)
        assert abs(val.compute_mmd(X, Y) - float(g[f"mmd_{tag}"])) < 1e-4 * float(g[f"mmd_{tag}"])          # fp32 tolerance: rel 1e-4
        assert abs(val.compute_mmd(X, Y, gamma=0.5 / d) - float(g[f"mmd_{tag}_gamma2"])) < 1e-4 * float(g[f"mmd_{tag}_gamma2"])
        assert val.compute_mmd(X, X) == 0.0 and val.compute_mmd(X, X.copy()) == 0.0      # exactly 0.0, like the reference (identical cohorts are detected)
    bf = BiologicalValidator(CONFIG, precision="bf16")
    X, Y = _mmd_inputs(150, 120, 64, np.random.RandomState(11))
    assert abs(bf.compute_mmd(X, Y) - float(g["mmd_small"])) < 2e-2 * float(g["mmd_small"])


@pytest.mark.parametrize("n,m,d", [(1, 1, 8), (129, 300, 70), (1000, 777, 5142), (2048, 2048, 256)])
def test_mmd_matches_oracle_ragged_shapes(n, m, d):
    rs = np.random.RandomState(n + m)
    X = (rs.standard_normal((n, d)) * 0.9 + 4.0).astype(np.float32)      # un-centred, expression-scale offset
    Y = (rs.standard_normal((m, d)) * 1.2 + 4.1).astype(np.float32)
    ref = V.compute_mmd(X, Y)
    got = BiologicalValidator(CONFIG).compute_mmd(X, Y)
    assert abs(got - ref) < 1e-4 * max(ref, 1e-3)
    # the three Gram sums themselves
    from osteosarcoma_diffusionmodel_b200 import validation as val
    Xt, Yt = torch.from_numpy(X).cuda(), torch.from_numpy(Y).cuda()
    center = ((Xt.sum(0, dtype=torch.float64) + Yt.sum(0, dtype=torch.float64)) / (n + m)).float()
    sums = val._gram_partial_sums(Xt, Yt, 1.0 / d, center, (0, n), (0, m), 1).cpu().numpy()
    ref_sums = np.array(V.rbf_sums(X.astype(np.float64), Y.astype(np.float64), 1.0 / d))
    assert np.allclose(sums, ref_sums, rtol=2e-5)


def test_mmd_row_sharding_sums_to_the_whole():
    """Emulate 3 ranks on one GPU: partial sums over 128-aligned row shards add up to the single-call result (§8e)."""
    from osteosarcoma_diffusionmodel_b200 import distributed as D, validation as val
    rs = np.random.RandomState(7)
    n, m, d = 700, 450, 96
    X = torch.from_numpy(rs.standard_normal((n, d)).astype(np.float32)).cuda()
    Y = torch.from_numpy((rs.standard_normal((m, d)) + 0.2).astype(np.float32)).cuda()
    center = ((X.sum(0) + Y.sum(0)) / (n + m)).contiguous()
    whole = val._gram_partial_sums(X, Y, 1.0 / d, center, (0, n), (0, m), 1)
    parts = sum(val._gram_partial_sums(X, Y, 1.0 / d, center, D.shard_rows(n, r, 3, 128), D.shard_rows(m, r, 3, 128), 1) for r in range(3))
    assert torch.allclose(whole, parts, rtol=1e-6)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_mmd_block_cyclic_symmetric_sharding_sums_to_the_whole(world):
    """BASELINE.json configs[4] sharding, emulated on one GPU: rank r reduces the 128-row blocks b % world == r of each Gram, K(X,X) and
    K(Y,Y) as symmetric half-Grams; the partial sums of all ranks equal the single-call sums (ragged last blocks, more ranks than blocks
    of Y included)."""
    from osteosarcoma_diffusionmodel_b200 import validation as val
    rs = np.random.RandomState(17)
    n, m, d = 1100, 300, 96
    X = torch.from_numpy(rs.standard_normal((n, d)).astype(np.float32)).cuda()
    Y = torch.from_numpy((rs.standard_normal((m, d)) + 0.2).astype(np.float32)).cuda()
    center = ((X.sum(0) + Y.sum(0)) / (n + m)).contiguous()
    for prec in (0, 1):
        whole = val._gram_partial_sums(X, Y, 1.0 / d, center, (0, n), (0, m), prec)
        parts = sum(val._gram_partial_sums_cyclic(X, Y, 1.0 / d, center, r, world, prec) for r in range(world))
        assert torch.allclose(whole, parts, rtol=1e-6), (prec, whole, parts)
    ref = np.array(V.rbf_sums(X.cpu().numpy().astype(np.float64), Y.cpu().numpy().astype(np.float64), 1.0 / d))
    assert np.allclose(parts.cpu().numpy(), ref, rtol=2e-5)


def test_pathway_coherence_matches_reference_golden(golden_dir):
    import pandas as pd
    g = np.load(golden_dir / "validators.npz")
    real, syn, members = coherence_inputs(g)
    n_genes = real.shape[1]
    genes = [f"G{i}" for i in range(n_genes)]
    gpm = pd.DataFrame(0, index=genes, columns=[f"P{p}" for p in range(12)])
    for p, idx in enumerate(members):
        gpm.iloc[idx, p] = 1
    gpm.iloc[[0, 1], 10] = 1      # an 11th pathway that must be ignored (only the first 10 count)
    val = BiologicalValidator(CONFIG)
    res = val.validate_pathway_coherence(pd.DataFrame(real, columns=genes), pd.DataFrame(syn, columns=genes), gpm)
    for k in ("real_pathway_coherence", "synthetic_pathway_coherence", "pathway_coherence_correlation"):
        assert abs(res[k] - float(g[f"coh_{k}"])) < 1e-6, k
    res2 = val.pathway_coherence_from_tensors(torch.from_numpy(real), torch.from_numpy(syn), members)
    assert res2 == pytest.approx(res, abs=1e-12)
    assert val.validate_pathway_coherence(pd.DataFrame(real[:, :2], columns=genes[:2]), pd.DataFrame(syn[:, :2], columns=genes[:2]), gpm) == {}


def test_mutation_expression_matches_reference_golden(golden_dir):
    import pandas as pd
    g = np.load(golden_dir / "validators.npz")
    mut, path = mutexpr_inputs()
    mdf = pd.DataFrame(mut, columns=["TP53", "MYC"])
    pdf = pd.DataFrame(path, columns=["HALLMARK_P53_PATHWAY", "HALLMARK_MYC_TARGETS_V1"])
    val = BiologicalValidator(CONFIG)
    res = val.validate_mutation_expression_correlation(mdf, pd.DataFrame(), pdf)
    assert res["mutation_expression_violation_rate"] == float(g["mutexpr_violation_rate"])
    assert val.validate_mutation_expression_correlation(mdf[["MYC"]], pd.DataFrame(), pdf[["HALLMARK_P53_PATHWAY"]]) == {}


def test_moments_large_cohort_against_numpy():
    from osteosarcoma_diffusionmodel_b200 import validation as val
    rs = np.random.RandomState(3)
    n = 200_000
    data = (rs.standard_normal((n, 40)) * 2 + 6).astype(np.float32)
    data[:, 5] = 0.7 * data[:, 3] + 0.3 * data[:, 5]
    cols = [3, 5, 17, 39, 0]
    t = torch.from_numpy(data).cuda()
    mom = val._moments(t, cols, t[0, cols].contiguous(), (0, n)).cpu().numpy()
    corr = val._corr_from_moments(mom, len(cols))
    ref = np.corrcoef(data[:, cols].astype(np.float64), rowvar=False)
    assert np.abs(corr - ref).max() < 1e-9


@pytest.mark.parametrize("n,ncols,pitch", [(200_000, 371, 371), (1001, 50, 64), (63, 33, 33)])
def test_batched_moments_match_numpy_and_the_single_set_kernel(n, ncols, pitch):
    """osteo_corr_moments_batched (one pass, warp per column set) against numpy and against the per-set kernel: set sizes 1 / 15 / 17 /
    32, a padded row pitch (scalar staging path), ragged last chunk, and row sharding (partial sums add up)."""
    from osteosarcoma_diffusionmodel_b200 import _lib, validation as val
    rs = np.random.RandomState(11)
    full = (rs.standard_normal((n, pitch)) * 1.5 + 4).astype(np.float32)
    full[:, 7] = 0.5 * full[:, 2] + 0.5 * full[:, 7]
    t = torch.from_numpy(full).cuda()
    sets = [sorted(rs.choice(ncols, k, replace=False).tolist()) for k in (1, 15, 17, 32)]
    if n < 2000:
        sets = sets + [sorted(rs.choice(ncols, 3 + i % 5, replace=False).tolist()) for i in range(28)]      # 32 sets: two launches of 16 warps
    ci = np.full((len(sets), 32), -1, dtype=np.int32)
    for i, c in enumerate(sets):
        ci[i, :len(c)] = c
    ci_t = torch.from_numpy(ci).cuda()
    shift = t[0, ci_t.clamp(min=0).long()].contiguous()
    lib, s = _lib.load(), _lib.stream_handle()

    def run(rb, re):
        out = torch.empty((len(sets), val._CM_STRIDE), dtype=torch.float64, device="cuda")
        _lib.check(lib.osteo_corr_moments_batched(t.data_ptr(), n, pitch, ncols, ci_t.data_ptr(), len(sets), shift.data_ptr(), rb, re, out.data_ptr(), s))
        return out.cpu().numpy()

    whole = run(0, n)
    cut = (n // 3 // 8) * 8 + 5
    parts = run(0, cut) + run(cut, n)
    assert np.allclose(whole, parts, rtol=1e-6, atol=1e-4)
    for i, cols in enumerate(sets):
        k = len(cols)
        blk = whole[i]
        assert blk[0] == n
        mom = np.concatenate([blk[:1], blk[1:1 + k], blk[33:].reshape(32, 32)[:k, :k].reshape(-1)])
        single = val._moments(t, cols, shift[i, :k].contiguous(), (0, n)).cpu().numpy()
        assert np.allclose(mom, single, rtol=2e-6, atol=1e-3)      # fp32 per-chunk partial sums vs fp64 products
        if k > 1:
            corr = val._corr_from_moments(mom, k)
            ref = np.corrcoef(full[:, cols].astype(np.float64), rowvar=False)
            assert np.abs(corr - ref).max() < 1e-6


@pytest.mark.parametrize("n,ncols,pitch,sizes", [(1000, 64, 64, [3, 15, 16, 1, 7]), (50_001, 371, 371, [15] * 10), (4099, 371, 371, [11, 19, 17, 16, 32, 3, 12]),
                                                 (777, 40, 56, [9, 32]), (31, 20, 20, [5, 17])])
def test_tiled_moments_and_device_finish_match_numpy(n, ncols, pitch, sizes):
    """osteo_corr_moments_tiled (4 x 4 register blocks; the 10-block form for sets <= 16 columns, the 36-block form up to 32, several
    passes when the sets do not fit one launch) + osteo_coherence_finish against float64 numpy: moment blocks, Pearson matrices and the
    per-set coherence score; ragged last chunk, row pitch > columns, a row sub-range."""
    from osteosarcoma_diffusionmodel_b200 import validation as val
    rs = np.random.RandomState(n + ncols)
    base = rs.standard_normal((n, 4)) @ rs.standard_normal((4, ncols)) * 0.5 + rs.standard_normal((n, ncols)) + 3.0
    buf = np.zeros((n, pitch), dtype=np.float32)
    buf[:, :ncols] = base.astype(np.float32)
    t = torch.from_numpy(buf).cuda()[:, :ncols]
    sets = [sorted(rs.choice(ncols, k, replace=False).tolist()) for k in sizes]
    ci = np.full((len(sets), 32), -1, dtype=np.int32)
    for i, c in enumerate(sets):
        ci[i, :len(c)] = c
    ci_t = torch.from_numpy(ci).cuda()
    rb, re = (0, n) if n < 1000 else (7, n - 5)
    mom = val._moments_tiled(t, ci_t, (rb, re), max(sizes))
    scores = val._coherence_finish(mom, ci_t).cpu().numpy()
    mom = mom.cpu().numpy()
    x64 = buf[:, :ncols].astype(np.float64)
    for i, c in enumerate(sets):
        k = len(c)
        g = x64[rb:re][:, c] - x64[0, c]
        assert mom[i, 0] == re - rb
        assert np.allclose(mom[i, 1:1 + k], g.sum(0), rtol=1e-6, atol=1e-3)
        s2 = mom[i, 33:].reshape(32, 32)[:k, :k]
        assert np.allclose(s2, g.T @ g, rtol=2e-6, atol=1e-3)
        assert np.array_equal(s2, s2.T)
        if k >= 2:
            ref = np.corrcoef(x64[rb:re][:, c], rowvar=False)
            got = val._corr_from_moments(np.concatenate([mom[i, :1], mom[i, 1:1 + k], s2.reshape(-1)]), k)
            assert np.abs(got - ref).max() < 1e-6
            assert abs(scores[i] - ref[np.triu_indices(k, k=1)].mean()) < 1e-6
