"""CPU: numpy restatement of the validators against the reference's outputs (tests/golden/validators.npz)."""
import numpy as np

from oracle import validators_oracle as V


def _mmd_inputs(n, m, d, rs):
    X = rs.standard_normal((n, d)) * 1.0 + 0.3
    Y = rs.standard_normal((m, d)) * 1.1
    return X, Y


def test_mmd_matches_reference(golden_dir):
    g = np.load(golden_dir / "validators.npz")
    rs = np.random.RandomState(11)
    for tag in ("small", "wide"):
        n, m, d = (int(v) for v in g[f"mmd_{tag}_shape"])
        X, Y = _mmd_inputs(n, m, d, rs)
        assert abs(V.compute_mmd(X, Y) - float(g[f"mmd_{tag}"])) < 1e-12
        assert abs(V.compute_mmd(X, Y, gamma=0.5 / d) - float(g[f"mmd_{tag}_gamma2"])) < 1e-12
        assert V.compute_mmd(X, X) == float(g[f"mmd_{tag}_self"]) == 0.0


def test_sqeuclidean_expansion_agrees_with_literal():
    rs = np.random.RandomState(0)
    X = rs.standard_normal((37, 19)) + 2.0
    Y = rs.standard_normal((23, 19)) - 1.0
    diff = X[:, None, :] - Y[None, :, :]
    lit = (diff * diff).sum(-1)
    nx, ny = (X * X).sum(1), (Y * Y).sum(1)
    exp = np.maximum(nx[:, None] + ny[None, :] - 2 * X @ Y.T, 0)
    assert np.allclose(V.sqeuclidean(X, Y), lit, rtol=0, atol=1e-12)
    assert np.allclose(exp, lit, rtol=1e-12, atol=1e-11)


def coherence_inputs(g):
    n_genes = int(g["coh_n_genes"])
    rs = np.random.RandomState(12)
    load = 0.45 + 0.35 * rs.standard_normal((n_genes, 6))
    real = rs.standard_normal((300, 6)) @ load.T + rs.standard_normal((300, n_genes))
    syn = rs.standard_normal((260, 6)) @ (load * 0.8).T + rs.standard_normal((260, n_genes)) * 1.1
    lens = g["coh_member_len"]
    flat = g["coh_member_idx"]
    members, o = [], 0
    for l in lens:
        members.append([int(v) for v in flat[o:o + l]])
        o += int(l)
    return real.astype(np.float32), syn.astype(np.float32), members


def test_pathway_coherence_matches_reference(golden_dir):
    g = np.load(golden_dir / "validators.npz")
    real, syn, members = coherence_inputs(g)
    assert len(members) == 10 and tuple(g["gpm_shape"]) == (371, 29)
    res = V.pathway_coherence(real, syn, members)
    for k in ("real_pathway_coherence", "synthetic_pathway_coherence", "pathway_coherence_correlation"):
        assert abs(res[k] - float(g[f"coh_{k}"])) < 1e-10, k
    assert res["real_pathway_coherence"] > 0.2


def test_pathway_coherence_skips_small_sets():
    rs = np.random.RandomState(1)
    a, b = rs.standard_normal((50, 8)), rs.standard_normal((40, 8))
    assert V.pathway_coherence(a, b, [[0, 1]]) == {}
    assert "real_pathway_coherence" in V.pathway_coherence(a, b, [[0, 1], [2, 3, 4]])


def mutexpr_inputs():
    rs = np.random.RandomState(13)
    n = 400
    mut = (rs.random_sample((n, 2)) < 0.3).astype(np.float64)
    path = rs.standard_normal((n, 2))
    path[:, 0] -= 0.8 * mut[:, 0]
    path[:, 1] -= 0.5 * mut[:, 1]
    return mut, path


def test_mutation_expression_matches_reference(golden_dir):
    g = np.load(golden_dir / "validators.npz")
    mut, path = mutexpr_inputs()
    res = V.mutation_expression_violation_rate([(mut[:, 0], path[:, 0], "negative"), (mut[:, 1], path[:, 1], "positive")])
    assert res["mutation_expression_violation_rate"] == float(g["mutexpr_violation_rate"]) == 0.5
    assert np.allclose(res["_correlations"], g["mutexpr_corr"], rtol=0, atol=1e-12)


def test_bio_loss_oracle_is_tied_to_the_validator_oracle():
    """oracle/bio_losses_oracle.py (the checker of the A12 kernels): coherence loss = 1 - per-pathway validator score, rule loss > 0
    exactly on a sign violation; its autograd gradient matches central differences."""
    import torch
    from oracle import bio_losses_oracle as B

    rs = np.random.RandomState(2)
    base = rs.standard_normal((200, 3))
    data = base @ rs.standard_normal((3, 12)) + 0.5 * rs.standard_normal((200, 12))
    sets = [[0, 1, 2, 3], [4, 5, 6, 7, 8], [9, 10], [9, 11]]
    modes = [0, 0, 1, -1]
    x = torch.from_numpy(data).requires_grad_(True)
    losses = B.correlation_losses(x, sets, modes)
    for s, m, l in zip(sets, modes, losses.tolist()):
        sub = data[:, s]
        if m == 0:
            assert abs(l - (1.0 - V.mean_upper(V.pearson_matrix(sub)))) < 1e-12
        else:
            c = V.pearson(sub[:, 0], sub[:, 1])
            assert abs(l - max(0.0, -m * c)) < 1e-12
            assert (l > 0) == ((m > 0 and c < 0) or (m < 0 and c > 0))
    w = torch.tensor([1.0, 0.5, 2.0, 3.0], dtype=torch.float64)
    (losses * w).sum().backward()
    for (r, c) in [(0, 0), (17, 5), (199, 9), (3, 11)]:
        h = 1e-6
        xp, xm = data.copy(), data.copy()
        xp[r, c] += h
        xm[r, c] -= h
        fd = ((B.correlation_losses(torch.from_numpy(xp), sets, modes) - B.correlation_losses(torch.from_numpy(xm), sets, modes)) * w).sum().item() / (2 * h)
        assert abs(fd - x.grad[r, c].item()) < 1e-6 * max(1.0, abs(fd))
