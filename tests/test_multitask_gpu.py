"""GPU: the differentiable biology losses and the multi-task training step against the torch-autograd oracle
(oracle/bio_losses_oracle.py, oracle/ddpm_oracle.py) and against the validators their forward values are tied to."""
import numpy as np
import pytest
import torch

from oracle import bio_losses_oracle as B
from oracle import ddpm_oracle as O
from oracle import synth
from oracle import validators_oracle as V
from osteosarcoma_diffusionmodel_b200.multitask import BiologyConstrainedDiffusion, correlation_losses
from tests.helpers import load_case, oracle_sd, rel

pytestmark = pytest.mark.gpu


def _cohort(n, g, seed):
    rs = np.random.RandomState(seed)
    base = rs.standard_normal((n, 4))
    mix = rs.standard_normal((4, g))
    return torch.from_numpy((base @ mix + 0.7 * rs.standard_normal((n, g)) + rs.standard_normal(g) * 3).astype(np.float32))


@pytest.mark.parametrize("n", [7, 500, 4099])
def test_correlation_losses_match_oracle_values_and_gradients(n):
    g = 70
    data = _cohort(n, g, 3)
    data[:, 1] = (data[:, 1] > data[:, 1].median()).float()       # a 0/1 mutation column
    sets = [[0, 5, 9, 11, 12], list(range(20, 52)), [3, 4, 60], [1, 30], [1, 31], [2, 69]]
    modes = [0, 0, 0, 1, -1, -1]
    x = data.cuda().requires_grad_(True)
    losses = correlation_losses(x, sets, modes)
    w = torch.linspace(0.5, 1.5, len(sets), device="cuda")
    (losses * w).sum().backward()
    xr = data.double().requires_grad_(True)
    ref = B.correlation_losses(xr, sets, modes)
    (ref * w.cpu().double()).sum().backward()
    assert torch.allclose(losses.cpu().double(), ref.detach(), atol=2e-6, rtol=1e-5)
    assert rel(x.grad, xr.grad) < 2e-5
    # forward values tied to the validators (utils/validation.py:150-157, :206-214)
    for s, m, l in zip(sets, modes, losses.tolist()):
        sub = data.numpy().astype(np.float64)[:, s]
        if m == 0:
            assert abs(l - (1.0 - V.mean_upper(V.pearson_matrix(sub)))) < 1e-5
        else:
            corr = V.pearson(sub[:, 0], sub[:, 1])
            violation = (m > 0 and corr < 0) or (m < 0 and corr > 0)
            assert (l > 0) == violation and abs(l - max(0.0, -m * corr)) < 1e-5


def test_more_than_32_sets_and_argument_checks():
    data = _cohort(300, 40, 1).cuda()
    sets = [[i, (i + 1) % 40, (i + 7) % 40] for i in range(40)]
    losses = correlation_losses(data, sets, [0] * 40)
    ref = B.correlation_losses(data.cpu(), sets, [0] * 40)
    assert torch.allclose(losses.cpu().double(), ref, atol=2e-6)
    with pytest.raises(ValueError):
        correlation_losses(data, [[0]], [0])
    with pytest.raises(ValueError):
        correlation_losses(data, [[0, 1, 2]], [1])
    with pytest.raises(RuntimeError):
        correlation_losses(data.cpu(), [[0, 1]], [0])


def _wrapper(case, members, rules, precision="fp32x3"):
    cfg = synth.model_config(hidden_dims=case["hidden"], schedule=case["schedule"], num_steps=case["T"], dropout=0.2)
    cfg["model"]["constraints"] = {"pathway_coherence_weight": 1.0, "mutation_expression_weight": 0.5, "survival_prediction_weight": 0.3}
    d = case["dims"]
    torch.manual_seed(4)
    m = BiologyConstrainedDiffusion(d["mutation_dim"], d["expression_dim"], d["pathway_dim"], d["condition_dim"], cfg, pathway_members=members,
                                    correlation_rules=rules)
    m.diffusion.load_state_dict(case["sd"], strict=False)
    m = m.cuda()
    m.diffusion.set_precision(precision)
    m.survival_predictor[2].p = 0.0        # the head's Dropout draws from torch's RNG: off for the parity run
    return m.train()


def test_multitask_step_matches_oracle_autograd():
    """loss parts and EVERY parameter gradient (denoiser + survival head) of one multi-task step, same injected t / noise /
    dropout masks on both sides."""
    case = load_case("linear3")
    sd = oracle_sd(case)
    T, D, rows = case["T"], case["D"], 300
    d = case["dims"]
    M, E, P = d["mutation_dim"], d["expression_dim"], d["pathway_dim"]
    x0, cond = synth.make_cohort(rows, M, E, P, d["condition_dim"], seed=9)
    rs = np.random.RandomState(5)
    t = torch.from_numpy(rs.randint(0, T, size=rows).astype(np.int64))
    t[:40] = torch.from_numpy(rs.randint(0, 30, size=40))           # enough rows with signal (ab_t >= 0.5)
    noise = synth.noise_stream(21)(1, (rows, D))
    masks = synth.dropout_masks(21, rows, synth.block_widths(case["hidden"]), 0.2)
    survival = torch.from_numpy(rs.standard_normal(rows).astype(np.float32))
    members = [[0, 3, 5, 8], [10, 11, 12, 13, 14, 15, 16], [3, 40, 41]]
    rules = [(2, 1, -1), (4, 0, 1), (5, 1, 1)]
    model = _wrapper(case, members, rules)
    model.diffusion._inject = {"t": t, "noise": noise, "masks": masks}
    total = model(x0.cuda(), cond.cuda(), survival_time=survival.cuda())
    total.backward()

    # ---- oracle: torch autograd over the CPU restatement
    pnames = [n for n, _ in synth.param_shapes(D, d["condition_dim"], case["hidden"])]
    params = {k: sd[k].clone().requires_grad_(True) for k in pnames}
    full = dict(sd)
    full.update(params)
    head = [p.detach().cpu().clone().requires_grad_(True) for p in model.survival_predictor.parameters()]
    x_t = O.q_sample(full, x0, t, noise)
    eps = O.predict_eps(full, x_t, t, cond, T, drop_masks=masks, p=0.2, training=True)
    mse = torch.nn.functional.mse_loss(eps, noise)
    ab = full["alphas_cumprod"][t]
    x0hat = (x_t - torch.sqrt(1 - ab)[:, None] * eps) / torch.sqrt(ab)[:, None]
    keep = (ab >= 0.5).nonzero().squeeze(1)
    assert keep.numel() >= 40
    sub = x0hat[keep]
    sets = [[M + g for g in m] for m in members] + [[mc, M + E + pc] for mc, pc, _ in rules]
    modes = [0] * len(members) + [s for _, _, s in rules]
    cl = B.correlation_losses(sub, sets, modes).float()
    u = torch.cat([sub[:, :M], sub[:, M + E:]], dim=1)[:, :128]
    u = torch.nn.functional.pad(u, (0, 128 - u.shape[1]))
    h = torch.relu(u @ head[0].t() + head[1])
    pred = (h @ head[2].t() + head[3]).squeeze(-1)
    surv = torch.nn.functional.mse_loss(pred, survival[keep])
    ref_total = mse + 1.0 * cl[:3].mean() + 0.5 * cl[3:].mean() + 0.3 * surv
    ref_total.backward()

    parts = {k: float(v) for k, v in model.last_losses.items()}
    assert abs(parts["diffusion"] - mse.item()) < 1e-5 * abs(mse.item())
    assert abs(parts["pathway_coherence"] - cl[:3].mean().item()) < 2e-4
    assert abs(parts["mutation_expression"] - cl[3:].mean().item()) < 2e-4
    assert abs(parts["survival"] - surv.item()) < 2e-4 * max(1.0, abs(surv.item()))
    assert abs(total.item() - ref_total.item()) < 2e-4 * abs(ref_total.item())
    for n, p in model.diffusion.named_parameters():
        assert rel(p.grad, params[n].grad) < 5e-4, n
    for p, r in zip(model.survival_predictor.parameters(), head):
        assert rel(p.grad, r.grad) < 5e-4
    model.diffusion.check_status()


def test_multitask_wrapper_trains_under_the_reference_recipe_in_bf16():
    """utils/train.py:230-246 on the wrapper, in-kernel noise / dropout, bf16 operands: finite, decreasing loss; eval forward and
    sample() delegate to the wrapped model."""
    case = load_case("smoke")
    d = case["dims"]
    model = _wrapper(case, [[0, 1, 2, 3], [5, 6, 7]], [(0, 0, -1)], precision="bf16")
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
    x0, cond = synth.make_cohort(256, d["mutation_dim"], d["expression_dim"], d["pathway_dim"], d["condition_dim"], seed=2)
    x0, cond = x0.cuda(), cond.cuda()
    surv = cond[:, 0].clone()
    first = last = None
    for i in range(12):
        opt.zero_grad()
        loss = model(x0, cond, survival_time=surv)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        assert torch.isfinite(loss)
        first = loss.item() if first is None else first
        last = loss.item()
    assert set(model.last_losses) == {"diffusion", "pathway_coherence", "mutation_expression", "survival"}
    assert last < first
    model.eval()
    with torch.no_grad():
        assert torch.isfinite(model(x0, cond))
    assert model.sample(cond[:8], 8, t_stop=990).shape == (8, case["D"])
    assert not hasattr(model, "vae")          # utils/train.py:233 dispatches on that attribute
    model.diffusion.check_status()


def test_correlation_losses_of_an_empty_selection_are_finite():
    """Every row filtered out (n_eff = 0: all timesteps too noisy for the auxiliary terms): the moment blocks are empty; the losses must be
    0 with zero gradient, not 0 / 0 (corr_loss_finish_kernel)."""
    import ctypes as C
    from osteosarcoma_diffusionmodel_b200 import _lib
    from osteosarcoma_diffusionmodel_b200.validation import _CM_STRIDE
    S = 3
    mom = torch.zeros((S, _CM_STRIDE), dtype=torch.float64, device="cuda")
    ci = torch.full((S, 32), -1, dtype=torch.int32, device="cuda")
    ci[0, :4] = torch.arange(4)
    ci[1, :2] = torch.tensor([0, 5])
    ci[2, :2] = torch.tensor([1, 6])
    shift = torch.zeros((S, 32), device="cuda")
    modes = torch.tensor([0, 1, -1], dtype=torch.int32, device="cuda")
    losses = torch.full((S,), float("nan"), device="cuda")
    coef = torch.full((S * 32 * 4,), float("nan"), device="cuda")
    _lib.check(_lib.load().osteo_corr_loss_finish(mom.data_ptr(), ci.data_ptr(), shift.data_ptr(), modes.data_ptr(), S, losses.data_ptr(), coef.data_ptr(), _lib.stream_handle()))
    assert torch.isfinite(losses).all() and float(losses.abs().sum()) == 0.0
    assert torch.isfinite(coef).all()
    assert float(coef.view(S, 32, 4)[:, :, 2].abs().sum()) == 0.0        # d loss / dS coefficients: no gradient flows
