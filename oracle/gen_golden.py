"""ORACLE support — generate tests/golden/*.npz by RUNNING THE REFERENCE (build container only).

    python -m oracle.gen_golden

Imports the unmodified /root/reference (models/diffusion.py, utils/validation.py,
utils/pathway_features.py) under the torch_geometric stub, feeds it the deterministic inputs of
oracle/synth.py and injects every random draw (torch.randint / randn_like / randn / dropout), then
stores only the reference's OUTPUTS.  The fixtures travel with the repo; the reference does not.
"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import reference_import, synth  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
P_SAMPLE_STEPS = [999, 998, 997, 990, 900, 500, 100, 1, 0]


class Injector:
    """Patches torch.randint / randn_like / randn / F.dropout so the reference consumes given draws."""

    def __init__(self):
        self.randint_q, self.randn_like_q, self.randn_q, self.mask_q = [], [], [], []
        self.p = None

    def __enter__(self):
        import torch.nn.functional as F

        self._orig = (torch.randint, torch.randn_like, torch.randn, F.dropout)
        inj = self

        def randint(*a, **k):
            return inj.randint_q.pop(0).clone()

        def randn_like(x, *a, **k):
            v = inj.randn_like_q.pop(0)
            assert v.shape == x.shape
            return v.clone()

        def randn(*a, **k):
            return inj.randn_q.pop(0).clone()

        def dropout(x, p=0.5, training=True, inplace=False):
            if not training or p == 0.0:
                return x
            m = inj.mask_q.pop(0)
            assert m.shape == x.shape
            return x * m.to(x.dtype) / (1.0 - p)

        torch.randint, torch.randn_like, torch.randn, F.dropout = randint, randn_like, randn, dropout
        return self

    def __exit__(self, *exc):
        import torch.nn.functional as F

        torch.randint, torch.randn_like, torch.randn, F.dropout = self._orig


def build_reference_model(ref_diffusion, dims, hidden, schedule, seed):
    cfg = synth.model_config(hidden_dims=hidden, schedule=schedule)
    model = ref_diffusion.BiologyAwareDiffusionModel(dims["mutation_dim"], dims["expression_dim"], dims["pathway_dim"], dims["condition_dim"], cfg)
    D = dims["mutation_dim"] + dims["expression_dim"] + dims["pathway_dim"]
    sd = synth.make_params(D, dims["condition_dim"], hidden, seed=seed)
    missing = model.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys and set(missing.missing_keys) <= {"betas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"}, missing
    return model, cfg


def subsample(t: torch.Tensor, limit: int = 512) -> np.ndarray:
    f = t.detach().reshape(-1)
    stride = max(1, f.numel() // limit)
    return f[::stride][:limit].numpy().copy()


def gen_ddpm_case(ref_diffusion, name, dims, hidden, schedule, batch, seed, full_loop_rows):
    torch.manual_seed(0)
    model, cfg = build_reference_model(ref_diffusion, dims, hidden, schedule, seed)
    D = model.data_dim
    T = model.num_steps
    x0, cond = synth.make_cohort(batch, dims["mutation_dim"], dims["expression_dim"], dims["pathway_dim"], dims["condition_dim"], seed=seed)
    draw = synth.noise_stream(seed)
    out = {"data_dim": D, "num_steps": T, "batch": batch, "seed": seed, "hidden": np.array(hidden)}
    for k in ("betas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"):
        out[f"buf_{k}"] = getattr(model, k).numpy().copy()
    # TimeEmbedding rows as p_sample feeds it
    temb_rows = [0, 1, 2, 3, 250, 499, 500, 750, 998, 999]
    tn = torch.tensor([t / T for t in temb_rows], dtype=torch.float32)
    out["temb_rows"] = np.array(temb_rows)
    out["temb"] = model.unet.time_embed(tn).numpy().copy()

    # ---- eval forward: eps_hat for injected t / noise (models/diffusion.py:344-380)
    rs = np.random.RandomState(seed + 5)
    t_idx = torch.from_numpy(rs.randint(0, T, size=batch).astype(np.int64))
    t_idx[0] = 0
    t_idx[-1] = T - 1
    noise = draw(1, (batch, D))
    model.eval()
    with Injector() as inj, torch.no_grad():
        inj.randint_q.append(t_idx)
        inj.randn_like_q.append(noise)
        eps_hat = model(x0, cond, return_loss=False)
    with Injector() as inj, torch.no_grad():
        inj.randint_q.append(t_idx)
        inj.randn_like_q.append(noise)
        loss_eval = model(x0, cond, return_loss=True)
    x_t, _ = model.q_sample(x0, t_idx, noise)
    out.update(t_idx=t_idx.numpy(), eps_hat=eps_hat.numpy().copy(), loss_eval=np.float32(loss_eval.item()), q_sample=x_t.numpy().copy())

    # ---- train forward + backward with injected dropout masks (utils/train.py:236-239)
    masks = synth.dropout_masks(seed, batch, synth.block_widths(hidden), cfg["model"]["gnn"]["dropout"])
    model.train()
    model.zero_grad()
    with Injector() as inj:
        inj.randint_q.append(t_idx)
        inj.randn_like_q.append(noise)
        inj.mask_q.extend(masks)
        loss_train = model(x0, cond, return_loss=True)
        assert not inj.mask_q
    loss_train.backward()
    out["loss_train"] = np.float32(loss_train.item())
    names = [n for n, _ in model.named_parameters()]
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array([p.grad.double().norm().item() for _, p in model.named_parameters()])
    for i, (n, p) in enumerate(model.named_parameters()):
        out[f"grad_sub_{i}"] = subsample(p.grad)
    model.eval()

    # ---- single reverse steps from a common x_t (models/diffusion.py:382-425)
    x_start = draw(2, (batch, D)) * 1.5
    out["p_sample_steps"] = np.array(P_SAMPLE_STEPS)
    eps_list, next_list = [], []
    for t in P_SAMPLE_STEPS:
        z = draw(100 + t, (batch, D))
        # eps via the unet directly (same call p_sample makes)
        with torch.no_grad():
            tn1 = torch.full((batch,), t / T)
            eps = model.unet(x_start, tn1, model.condition_embed(cond))
        with Injector() as inj:
            if t > 0:
                inj.randn_like_q.append(z)
            nxt = model.p_sample(x_start, t, cond)
        eps_list.append(eps.numpy().copy())
        next_list.append(nxt.numpy().copy())
    out["p_sample_eps"] = np.stack(eps_list)
    out["p_sample_next"] = np.stack(next_list)

    # ---- the full 1000-step loop (models/diffusion.py:427-449) with injected x_T and per-step z
    rows = full_loop_rows
    cond_loop = synth.scenario_conditions(rows, dims["condition_dim"]) if dims["condition_dim"] == 3 else cond[:rows]
    with Injector() as inj:
        inj.randn_q.append(draw(3, (rows, D)))
        for t in reversed(range(1, T)):
            inj.randn_like_q.append(draw(10_000 + t, (rows, D)))
        final = model.sample(cond_loop, num_samples=rows)
        assert not inj.randn_like_q
    out["loop_rows"] = rows
    out["loop_final"] = final.numpy().copy()
    # intermediate checkpoints of the same loop (re-run with the oracle-independent reference p_sample)
    x = draw(3, (rows, D))
    ck_steps = [990, 900, 500, 100, 10]
    cks = []
    for t in reversed(range(T)):
        with Injector() as inj:
            if t > 0:
                inj.randn_like_q.append(draw(10_000 + t, (rows, D)))
            x = model.p_sample(x, t, cond_loop)
        if t in ck_steps:
            cks.append(x.numpy().copy())
    assert np.array_equal(x.numpy(), out["loop_final"])
    out["loop_ck_steps"] = np.array(ck_steps)
    out["loop_ck"] = np.stack(cks)
    GOLDEN.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLDEN / f"ddpm_{name}.npz", **out)
    print(f"wrote ddpm_{name}.npz  loss_eval={loss_eval.item():.6f} loss_train={loss_train.item():.6f} |final|={final.norm().item():.4e}")


def gen_production_case(ref_diffusion):
    """The reference's sample() on the 208 rows of oracle/production_case.py with x_T / z taken from the restated Philox streams: what
    the benchmarked mode (in-kernel RNG, graphs, branches) must reproduce for those rows of a 76 100-patient run."""
    from oracle import philox_oracle as P
    from oracle import production_case as PC

    torch.manual_seed(0)
    dims = synth.CONFIG_YAML_DIMS
    model, _ = build_reference_model(ref_diffusion, dims, (256, 512, 256), "cosine", PC.PARAM_SEED)
    model.eval()
    D, T = model.data_dim, model.num_steps
    rows = PC.rows()
    cond = synth.scenario_conditions(PC.N_TOTAL, 3)[torch.from_numpy(rows)]
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    with Injector() as inj:
        inj.randn_q.append(f32(P.normals(PC.SEED, rows.astype(np.uint64), D, PC.STREAM_XT, 0)))
        for t in reversed(range(1, T)):
            inj.randn_like_q.append(f32(P.normals(PC.SEED, rows.astype(np.uint64), D, PC.STREAM_REVERSE, t)))
        final = model.sample(cond, num_samples=len(rows))
        assert not inj.randn_like_q
    final = final.numpy()
    cols = PC.stored_columns(dims["mutation_dim"], dims["expression_dim"], dims["pathway_dim"])
    np.savez_compressed(GOLDEN / "ddpm_production.npz", rows=rows, n_total=PC.N_TOTAL, seed=PC.SEED, cols=cols, final_cols=final[:, cols],
                        final_norm=np.float64(np.linalg.norm(final.astype(np.float64))), final_absmax=np.float32(np.abs(final).max()))
    print(f"wrote ddpm_production.npz rows={len(rows)} |final|={np.linalg.norm(final):.4e} max|x|={np.abs(final).max():.4e}")


def gen_validators(ref_validation):
    import pandas as pd
    import importlib.util

    spec = importlib.util.spec_from_file_location("_ref_pathway_features", os.path.join(reference_import.REFERENCE_ROOT, "utils", "pathway_features.py"))
    pf = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pf)
    eng = pf.PathwayFeatureEngineering()
    eng.load_gene_sets()
    gpm = eng.create_gene_pathway_matrix()          # [371 genes, 29 pathways] 0/1
    genes = list(gpm.index)
    config = {"evaluation": {"driver_genes": ["TP53", "RB1"], "mutually_exclusive_pairs": [["TP53", "MDM2"]],
                             "required_correlations": [{"mutation": "TP53", "pathway": "HALLMARK_P53_PATHWAY", "direction": "negative"},
                                                       {"mutation": "MYC", "pathway": "HALLMARK_MYC_TARGETS_V1", "direction": "positive"}]}}
    val = ref_validation.BiologicalValidator(config)
    out = {}
    # --- MMD (utils/validation.py:273-298)
    rs = np.random.RandomState(11)
    for tag, (n, m, d) in {"small": (150, 120, 64), "wide": (48, 40, 5142)}.items():
        X = rs.standard_normal((n, d)) * 1.0 + 0.3
        Y = rs.standard_normal((m, d)) * 1.1
        out[f"mmd_{tag}_shape"] = np.array([n, m, d])
        out[f"mmd_{tag}"] = np.float64(val.compute_mmd(X, Y))
        out[f"mmd_{tag}_gamma2"] = np.float64(val.compute_mmd(X, Y, gamma=0.5 / d))
        out[f"mmd_{tag}_self"] = np.float64(val.compute_mmd(X, X))
    # --- pathway coherence (utils/validation.py:125-175) on the reference's own gene x pathway matrix
    n_real, n_syn = 300, 260
    rs = np.random.RandomState(12)
    load = 0.45 + 0.35 * rs.standard_normal((len(genes), 6))
    real = rs.standard_normal((n_real, 6)) @ load.T + rs.standard_normal((n_real, len(genes)))
    syn = rs.standard_normal((n_syn, 6)) @ (load * 0.8).T + rs.standard_normal((n_syn, len(genes))) * 1.1
    real_df = pd.DataFrame(real.astype(np.float32), columns=genes)
    syn_df = pd.DataFrame(syn.astype(np.float32), columns=genes)
    coh = val.validate_pathway_coherence(real_df, syn_df, gpm)
    members = [[genes.index(g) for g in gpm[gpm[p] == 1].index] for p in gpm.columns[:10]]
    out["coh_n_genes"] = len(genes)
    out["coh_member_len"] = np.array([len(m) for m in members])
    out["coh_member_idx"] = np.concatenate([np.array(m) for m in members])
    for k, v in coh.items():
        out[f"coh_{k}"] = np.float64(v)
    out["gpm_shape"] = np.array(gpm.shape)
    # --- mutation vs pathway correlation (utils/validation.py:177-223)
    rs = np.random.RandomState(13)
    n = 400
    mut = (rs.random_sample((n, 2)) < 0.3).astype(np.float64)
    path = rs.standard_normal((n, 2))
    path[:, 0] -= 0.8 * mut[:, 0]      # TP53 mutated -> p53 pathway down (negative: no violation)
    path[:, 1] -= 0.5 * mut[:, 1]      # MYC: negative although 'positive' expected -> violation
    mdf = pd.DataFrame(mut, columns=["TP53", "MYC"])
    pdf = pd.DataFrame(path, columns=["HALLMARK_P53_PATHWAY", "HALLMARK_MYC_TARGETS_V1"])
    res = val.validate_mutation_expression_correlation(mdf, pd.DataFrame(), pdf)
    out["mutexpr_violation_rate"] = np.float64(res["mutation_expression_violation_rate"])
    out["mutexpr_corr"] = np.array([mdf["TP53"].corr(pdf["HALLMARK_P53_PATHWAY"]), mdf["MYC"].corr(pdf["HALLMARK_MYC_TARGETS_V1"])])
    np.savez_compressed(GOLDEN / "validators.npz", **out)
    print("wrote validators.npz", {k: float(v) for k, v in out.items() if np.ndim(v) == 0})


def gen_validate_all(ref_validation):
    """Outputs of the reference's remaining validators (utils/validation.py:27-123, :225-271, :300-387) on oracle/validator_inputs.py;
    the numpy global RNG is seeded before each call (np.random.choice of the 50 genes, sklearn's randomized PCA)."""
    from oracle import validator_inputs as VI

    val = ref_validation.BiologicalValidator(VI.CONFIG)
    out = {}
    real_mut, syn_mut = VI.mutation_frames()
    np.random.seed(123)
    for k, v in val.validate_mutation_cooccurrence(real_mut, syn_mut).items():
        out[f"cooc_{k}"] = np.float64(v)
    real, syn = VI.stat_matrices()
    np.random.seed(5)
    for k, v in val.statistical_tests(real, syn).items():
        out[f"stat_{k}"] = np.float64(v)
    # an asymptotic-branch KS case (n > 10 000): the statistic and p-value of the first 3 features
    from scipy import stats
    rs = np.random.RandomState(31)
    a, b = rs.standard_normal((12000, 3)), rs.standard_normal((10500, 3)) * 1.02 + 0.01
    ks = [stats.ks_2samp(a[:, i], b[:, i]) for i in range(3)]
    out["ks_big_stat"] = np.array([k.statistic for k in ks])
    out["ks_big_pvalue"] = np.array([k.pvalue for k in ks])
    frames = VI.validate_all_frames()
    np.random.seed(77)
    res = val.validate_all(*frames)
    out["all_keys"] = np.array(list(res.keys()))
    for k, v in res.items():
        out[f"all_{k}"] = np.float64(v)
    np.savez_compressed(GOLDEN / "validate_all.npz", **out)
    print("wrote validate_all.npz", {k: (float(v) if np.ndim(v) == 0 else v.tolist()) for k, v in out.items() if k != "all_keys"})


def main():
    ref_diffusion, ref_validation = reference_import.import_reference()
    torch.set_num_threads(os.cpu_count() or 1)
    GOLDEN.mkdir(parents=True, exist_ok=True)
    if "--only-validators" in sys.argv:
        gen_validators(ref_validation)
        gen_validate_all(ref_validation)
        return
    if "--only-production" in sys.argv:
        gen_production_case(ref_diffusion)
        return
    if "--only-validate-all" in sys.argv:
        gen_validate_all(ref_validation)
        return
    gen_ddpm_case(ref_diffusion, "smoke", synth.SMOKE_DIMS, (256, 512, 256), "cosine", batch=4, seed=1, full_loop_rows=4)
    gen_ddpm_case(ref_diffusion, "config", synth.CONFIG_YAML_DIMS, (256, 512, 256), "cosine", batch=4, seed=2, full_loop_rows=3)
    gen_ddpm_case(ref_diffusion, "linear3", dict(mutation_dim=20, expression_dim=90, pathway_dim=10, condition_dim=2), (128, 256), "linear", batch=5, seed=3,
                  full_loop_rows=2)
    gen_production_case(ref_diffusion)
    gen_validators(ref_validation)
    gen_validate_all(ref_validation)


if __name__ == "__main__":
    main()
