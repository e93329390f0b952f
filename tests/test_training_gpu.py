"""GPU: the training step (forward + fused backward) against the reference's own loss / gradients (golden fixtures)
and against oracle autograd on fresh inputs; plus the drop-in optimiser loop of utils/train.py:236-244."""
import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as O
from oracle import synth
from osteosarcoma_diffusionmodel_b200 import _lib
from tests.helpers import CASES, TOL_BF16, TOL_FP32X3, build_model, load_case, oracle_sd, rel

pytestmark = pytest.mark.gpu

TOL_GRAD_FP32X3 = 2e-4    # gradients chain ~25 contractions and the bf16x3 split drops the lo*lo terms
TOL_GRAD_BF16 = 5e-2      # bf16 operands AND bf16-stored x_hat / d(pre-norm) tensors


@pytest.mark.parametrize("rows,n_out,k_in", [(64, 128, 128), (1000, 256, 300), (77, 130, 70), (4096, 512, 1024), (300, 5142, 256)])
@pytest.mark.parametrize("prec,tol", [(_lib.PREC_FP32X3, 2e-5), (_lib.PREC_BF16, 6e-3)])
def test_wgrad_tc(rows, n_out, k_in, prec, tol):
    g = torch.Generator(device="cuda").manual_seed(3)
    dy = torch.randn(rows, n_out, device="cuda", generator=g)
    x = torch.randn(rows, k_in, device="cuda", generator=g)
    dw = torch.full((n_out, k_in), float("nan"), device="cuda")
    _lib.check(_lib.load().osteo_wgrad_tc(dy.data_ptr(), x.data_ptr(), dw.data_ptr(), rows, n_out, k_in, prec, None))
    ref = dy.double().t() @ x.double()
    assert ((dw.double() - ref).norm() / ref.norm()).item() < tol


def _inject(case, model, train=True):
    g = case["g"]
    noise = synth.noise_stream(case["seed"])(1, (case["batch"], case["D"]))
    masks = synth.dropout_masks(case["seed"], case["batch"], synth.block_widths(case["hidden"]), 0.2)
    model._inject = {"t": torch.from_numpy(g["t_idx"]), "noise": noise, "masks": masks if train else None}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("precision,tol_l,tol_g", [("fp32x3", 1e-5, TOL_GRAD_FP32X3), ("bf16", 5e-3, TOL_GRAD_BF16)])
def test_loss_and_gradients_match_reference(name, precision, tol_l, tol_g):
    case = load_case(name)
    g = case["g"]
    model = build_model(case, precision)
    x0, cond = case["x0"].cuda(), case["cond"].cuda()
    # eval-mode loss (utils/train.py:252-268 calls forward under no_grad + eval)
    _inject(case, model, train=False)
    with torch.no_grad():
        le = model(x0, cond)
    assert abs(le.item() - float(g["loss_eval"])) < tol_l * 10 * abs(float(g["loss_eval"]))
    # train-mode loss + backward with the same injected t / noise / dropout masks the reference consumed
    model.train()
    model.zero_grad()
    _inject(case, model, train=True)
    loss = model(x0, cond)
    loss.backward()
    assert abs(loss.item() - float(g["loss_train"])) < tol_l * 10 * abs(float(g["loss_train"]))
    names = [str(n) for n in g["grad_names"]]
    params = dict(model.named_parameters())
    assert names == list(params)
    worst = 0.0
    for i, n in enumerate(names):
        gr = params[n].grad
        assert gr is not None and torch.isfinite(gr).all(), n
        gn = gr.double().norm().item()
        ref_n = float(g["grad_norms"][i])
        assert abs(gn - ref_n) <= tol_g * max(ref_n, 1e-12), (n, gn, ref_n)
        f = gr.reshape(-1)
        stride = max(1, f.numel() // 512)
        e = rel(f[::stride][:512], g[f"grad_sub_{i}"])
        worst = max(worst, e)
        assert e < tol_g * 3, (n, e)      # sub-sampled entries: a few hundred values, looser than the full-norm bound
    model.check_status()


@pytest.mark.parametrize("rows", [300, 129])
def test_gradients_match_oracle_autograd_on_fresh_batch(rows):
    """Multi-tile, ragged batch, every parameter's full gradient against torch autograd over the oracle."""
    case = load_case("linear3")
    sd = oracle_sd(case)
    T, D = case["T"], case["D"]
    x0, cond = synth.make_cohort(rows, 20, 90, 10, 2, seed=9)
    rs = np.random.RandomState(5)
    t = torch.from_numpy(rs.randint(0, T, size=rows).astype(np.int64))
    noise = synth.noise_stream(21)(1, (rows, D))
    masks = synth.dropout_masks(21, rows, synth.block_widths(case["hidden"]), 0.2)
    pnames = [n for n, _ in synth.param_shapes(D, 2, case["hidden"])]
    params = {k: sd[k].clone().requires_grad_(True) for k in pnames}
    full = dict(sd)
    full.update(params)
    ref_loss = O.forward_loss(full, x0, cond, t, noise, T, drop_masks=masks, p=0.2, training=True)
    ref_loss.backward()
    model = build_model(case, "fp32x3")
    model.train()
    model._inject = {"t": t, "noise": noise, "masks": masks}
    loss = model(x0.cuda(), cond.cuda())
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 1e-5 * abs(ref_loss.item())
    for n, p in model.named_parameters():
        assert rel(p.grad, params[n].grad) < TOL_GRAD_FP32X3, n
    model.check_status()


def test_drop_in_optimizer_loop_tracks_the_oracle():
    """utils/train.py:230-246 verbatim (zero_grad, forward, backward, clip_grad_norm_(1.0), AdamW.step) for 3 steps, same
    injected draws on both sides: parameters stay within fp32x3 tolerance of torch autograd over the oracle."""
    case = load_case("linear3")
    T, D, rows = case["T"], case["D"], 64
    x0, cond = synth.make_cohort(rows, 20, 90, 10, 2, seed=2)
    model = build_model(case, "fp32x3")
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
    sd = oracle_sd(case)
    pnames = [n for n, _ in synth.param_shapes(D, 2, case["hidden"])]
    ref_params = {k: sd[k].clone().requires_grad_(True) for k in pnames}
    ref_opt = torch.optim.AdamW(list(ref_params.values()), lr=1e-3, weight_decay=1e-5)
    losses = []
    for step in range(3):
        rs = np.random.RandomState(100 + step)
        t = torch.from_numpy(rs.randint(0, T, size=rows).astype(np.int64))
        noise = synth.noise_stream(300 + step)(1, (rows, D))
        masks = synth.dropout_masks(300 + step, rows, synth.block_widths(case["hidden"]), 0.2)
        opt.zero_grad()
        model._inject = {"t": t, "noise": noise, "masks": masks}
        loss = model(x0.cuda(), cond.cuda(), return_loss=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        losses.append(loss.item())
        ref_opt.zero_grad()
        full = dict(sd)
        full.update(ref_params)
        rl = O.forward_loss(full, x0, cond, t, noise, T, drop_masks=masks, p=0.2, training=True)
        rl.backward()
        torch.nn.utils.clip_grad_norm_(list(ref_params.values()), 1.0)
        ref_opt.step()
        assert abs(loss.item() - rl.item()) < 1e-4 * abs(rl.item()), step
    for n, p in model.named_parameters():
        assert rel(p, ref_params[n]) < 1e-4, n
    model.check_status()


def test_in_kernel_draws_train_and_reduce_the_loss():
    """No injection: t from torch.randint, noise and dropout from the Philox streams; AdamW on a fixed batch lowers the loss."""
    case = load_case("linear3")
    rows = 256
    x0, cond = synth.make_cohort(rows, 20, 90, 10, 2, seed=4)
    x0, cond = x0.cuda(), cond.cuda()
    model = build_model(case, "bf16")
    model.train()
    model.manual_seed(3)
    opt = torch.optim.AdamW(model.parameters(), lr=2e-3)
    first = last = None
    for step in range(30):
        opt.zero_grad()
        loss = model(x0, cond)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        if step < 5:
            first = loss.item() if first is None else first + loss.item()
        if step >= 25:
            last = loss.item() if last is None else last + loss.item()
    assert np.isfinite(last) and last < first
    model.eval()
    with torch.no_grad():
        assert torch.isfinite(model(x0, cond))
    model.check_status()


def _train_run(case, graph: bool, steps: int = 5):
    """`steps` optimiser steps (utils/train.py:230-244) with in-kernel noise / dropout; returns losses, final parameters, last gradients."""
    model = build_model(case, "fp32x3")
    model.set_train_graph(graph)
    model.train()
    model.manual_seed(11)
    torch.manual_seed(11)            # the timestep draws (torch.randint, models/diffusion.py:361)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
    x0, cond = case["x0"].cuda(), case["cond"].cuda()
    losses = []
    for _ in range(steps):
        opt.zero_grad()
        loss = model(x0.clone(), cond.clone())      # fresh input addresses every step, like a DataLoader
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        losses.append(loss.item())
    model.check_status()
    return losses, [p.detach().clone() for p in model.parameters()], [p.grad.detach().clone() for p in model.parameters()]


@pytest.mark.parametrize("name", ["smoke", "linear3"])
def test_graph_replayed_training_matches_eager(name):
    """The step is replayed as one graph from its second call on (weights repack: from the third): same losses, gradients and
    parameters as the eager launches, and the in-kernel noise / dropout streams still advance from step to step."""
    case = load_case(name)
    l_g, p_g, g_g = _train_run(case, True)
    l_e, p_e, g_e = _train_run(case, False)
    assert len(set(l_g)) == len(l_g)                       # a frozen seed or timestep tensor would repeat a loss
    for a, b in zip(l_g, l_e):
        assert abs(a - b) < 2e-5 * abs(b)
    for a, b in zip(g_g, g_e):
        assert rel(a, b) < 1e-4 or (a - b).abs().max().item() < 1e-7      # split-batch wgrad accumulates atomically: not bit-stable
    for a, b in zip(p_g, p_e):
        assert rel(a, b) < 1e-3         # AdamW's g / sqrt(v) amplifies the atomics' rounding noise on near-zero gradients


def test_backward_after_a_second_forward_is_refused():
    case = load_case("smoke")
    model = build_model(case, "bf16")
    model.train()
    x0, cond = case["x0"].cuda(), case["cond"].cuda()
    first = model(x0, cond)
    second = model(x0, cond)
    with pytest.raises(RuntimeError, match="overwritten"):
        first.backward()
    second.backward()
    assert all(p.grad is not None for p in model.parameters())


@pytest.mark.parametrize("max_norm", [1.0, None])
def test_fused_adamw_tracks_torch_clip_and_adamw(max_norm):
    """osteo_adamw_step (two launches) against torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW.step (utils/train.py:242-244) on
    the same gradients for 6 steps: parameters, moments, clipped gradients and the reported norm; optimizer state_dicts are
    interchangeable."""
    from osteosarcoma_diffusionmodel_b200.optim import FusedAdamW

    g = torch.Generator(device="cuda").manual_seed(0)
    shapes = [(5142, 256), (256,), (512, 512), (3, 64), (1,), (4097,)]
    pa = [torch.nn.Parameter(torch.randn(s, device="cuda", generator=g)) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa = FusedAdamW(pa, lr=1e-2, weight_decay=1e-2, max_grad_norm=max_norm)
    ob = torch.optim.AdamW(pb, lr=1e-2, weight_decay=1e-2)
    for step in range(6):
        scale = 10.0 if step % 2 == 0 else 1e-3          # clipping active / inactive
        for a, b in zip(pa, pb):
            gr = torch.randn(a.shape, device="cuda", generator=g) * scale
            a.grad, b.grad = gr.clone(), gr.clone()
        v0 = [p._version for p in pa]
        oa.step()
        assert all(p._version > v for p, v in zip(pa, v0))       # weight caches keyed on the version see the update
        if max_norm is not None:
            ref_norm = torch.nn.utils.clip_grad_norm_(pb, max_norm)
            assert abs(oa.last_grad_norm.item() - ref_norm.item()) < 1e-5 * ref_norm.item()
            for a, b in zip(pa, pb):
                assert rel(a.grad, b.grad) < 1e-6
        ob.step()
        for a, b in zip(pa, pb):
            assert rel(a, b) < 2e-6
            assert rel(oa.state[a]["exp_avg"], ob.state[b]["exp_avg"]) < 2e-6
            assert rel(oa.state[a]["exp_avg_sq"], ob.state[b]["exp_avg_sq"]) < 2e-6
    # a torch AdamW checkpoint loads into the fused optimizer and vice versa
    oa.load_state_dict(ob.state_dict())
    ob.load_state_dict(oa.state_dict())
    assert int(oa.state[pa[0]]["step"]) == 6


def test_fused_adamw_drives_the_model():
    """The drop-in loop with the fused optimiser: the model repacks its weights after every step (version counters bumped) and the loss falls."""
    from osteosarcoma_diffusionmodel_b200.optim import FusedAdamW

    case = load_case("smoke")
    model = build_model(case, "bf16")
    model.train()
    opt = FusedAdamW(model.parameters(), lr=2e-3, weight_decay=1e-5, max_grad_norm=1.0)
    x0, cond = case["x0"].cuda(), case["cond"].cuda()
    torch.manual_seed(0)
    model.manual_seed(0)
    losses = []
    for _ in range(30):
        opt.zero_grad()
        loss = model(x0, cond)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < np.mean(losses[:5])
    model.check_status()


def test_graph_replay_survives_batch_changes_and_reallocation():
    """Alternating batch sizes (the larger one arrives late and forces the workspace to be re-allocated) and a sampling call in between:
    the cached graphs are keyed by shape, addresses and an allocation generation, so every step equals the eager step."""
    case = load_case("linear3")
    x0, cond = synth.make_cohort(700, 20, 90, 10, 2, seed=6)
    x0, cond = x0.cuda(), cond.cuda()
    order = [64, 64, 64, 300, 300, 64, 300, 700, 700, 64, 700]

    def run(graph):
        model = build_model(case, "fp32x3")
        model.set_train_graph(graph)
        model.train()
        model.manual_seed(3)
        torch.manual_seed(3)
        out = []
        for i, b in enumerate(order):
            model.zero_grad()
            loss = model(x0[:b], cond[:b])
            loss.backward()
            out.append((loss.item(), model.unet.output_proj.weight.grad.norm().item(), model.unet.input_proj.weight.grad.norm().item()))
            if i == 5:
                model.eval()
                model.sample(cond[:130], 130, seed=1, t_stop=995)
                model.train()
        model.check_status()
        return out

    for a, b in zip(run(True), run(False)):
        for u, v in zip(a, b):
            assert abs(u - v) < 2e-5 * abs(v)
