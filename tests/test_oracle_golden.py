"""CPU: the oracle restatement against fixtures produced by the reference itself (oracle/gen_golden.py).

Both sides execute the same ATen CPU kernels, so everything except the 1000-step loop (thread-count
dependent GEMM blocking) is compared at 1e-6 relative or tighter.
"""
import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as O
from oracle import synth

CASES = {
    "smoke": (synth.SMOKE_DIMS, (256, 512, 256), "cosine"),
    "config": (synth.CONFIG_YAML_DIMS, (256, 512, 256), "cosine"),
    "linear3": (dict(mutation_dim=20, expression_dim=90, pathway_dim=10, condition_dim=2), (128, 256), "linear"),
}


def load_case(golden_dir, name):
    g = np.load(golden_dir / f"ddpm_{name}.npz", allow_pickle=False)
    dims, hidden, schedule = CASES[name]
    D = dims["mutation_dim"] + dims["expression_dim"] + dims["pathway_dim"]
    seed = int(g["seed"])
    sd = synth.make_params(D, dims["condition_dim"], hidden, seed=seed)
    sd.update(O.schedule_buffers(schedule, int(g["num_steps"])))
    batch = int(g["batch"])
    x0, cond = synth.make_cohort(batch, dims["mutation_dim"], dims["expression_dim"], dims["pathway_dim"], dims["condition_dim"], seed=seed)
    return g, dims, hidden, schedule, D, seed, sd, x0, cond


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


@pytest.mark.parametrize("name", list(CASES))
def test_schedule_buffers_bit_exact(golden_dir, name):
    g, dims, hidden, schedule, D, seed, sd, x0, cond = load_case(golden_dir, name)
    for k in ("betas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"):
        assert np.array_equal(sd[k].numpy(), g[f"buf_{k}"]), k


def test_unknown_schedule_raises():
    with pytest.raises(ValueError, match="Unknown schedule"):
        O.beta_schedule("sigmoid", 10)


@pytest.mark.parametrize("name", list(CASES))
def test_time_embedding_rows(golden_dir, name):
    g, *_ = load_case(golden_dir, name)
    T = int(g["num_steps"])
    table = O.time_embedding_table(T, 128)
    assert np.array_equal(table[g["temb_rows"]].numpy(), g["temb"])
    # forward() normalises with t.float()/T, p_sample with python t/T: same fp32 for every t
    t = torch.arange(T)
    assert torch.equal(t.float() / T, torch.tensor([i / T for i in range(T)], dtype=torch.float32))


@pytest.mark.parametrize("name", list(CASES))
def test_q_sample_and_eps(golden_dir, name):
    g, dims, hidden, schedule, D, seed, sd, x0, cond = load_case(golden_dir, name)
    T = int(g["num_steps"])
    t = torch.from_numpy(g["t_idx"])
    noise = synth.noise_stream(seed)(1, (int(g["batch"]), D))
    xt = O.q_sample(sd, x0, t, noise)
    assert np.array_equal(xt.numpy(), g["q_sample"])
    with torch.no_grad():
        eps = O.predict_eps(sd, xt, t, cond, T)
        loss = O.forward_loss(sd, x0, cond, t, noise, T)
    assert rel(eps.numpy(), g["eps_hat"]) < 1e-6
    assert abs(loss.item() - float(g["loss_eval"])) < 1e-6 * abs(float(g["loss_eval"]))


@pytest.mark.parametrize("name", list(CASES))
def test_train_loss_and_grads(golden_dir, name):
    g, dims, hidden, schedule, D, seed, sd, x0, cond = load_case(golden_dir, name)
    T = int(g["num_steps"])
    t = torch.from_numpy(g["t_idx"])
    noise = synth.noise_stream(seed)(1, (int(g["batch"]), D))
    masks = synth.dropout_masks(seed, int(g["batch"]), synth.block_widths(hidden), 0.2)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k in dict(synth.param_shapes(D, dims["condition_dim"], hidden))}
    full = dict(sd)
    full.update(params)
    loss = O.forward_loss(full, x0, cond, t, noise, T, drop_masks=masks, p=0.2, training=True)
    loss.backward()
    assert abs(loss.item() - float(g["loss_train"])) < 2e-6 * abs(float(g["loss_train"]))
    names = [str(n) for n in g["grad_names"]]
    assert names == list(params.keys())
    for i, n in enumerate(names):
        gn = params[n].grad.double().norm().item()
        assert abs(gn - g["grad_norms"][i]) <= 1e-5 * max(g["grad_norms"][i], 1e-12), n
        f = params[n].grad.reshape(-1)
        stride = max(1, f.numel() // 512)
        sub = f[::stride][:512].numpy()
        assert rel(sub, g[f"grad_sub_{i}"]) < 1e-4, n


@pytest.mark.parametrize("name", list(CASES))
def test_p_sample_steps(golden_dir, name):
    g, dims, hidden, schedule, D, seed, sd, x0, cond = load_case(golden_dir, name)
    T = int(g["num_steps"])
    draw = synth.noise_stream(seed)
    B = int(g["batch"])
    x_start = draw(2, (B, D)) * 1.5
    cx, ce, sg = O.reverse_coefficients(sd["betas"], sd["alphas_cumprod"])
    for i, t in enumerate(g["p_sample_steps"]):
        t = int(t)
        z = draw(100 + t, (B, D))
        nxt, eps = O.p_sample(sd, x_start, t, cond, z if t > 0 else None, T, return_eps=True)
        assert rel(eps.numpy(), g["p_sample_eps"][i]) < 1e-6, t
        assert rel(nxt.numpy(), g["p_sample_next"][i]) < 1e-6, t
        # collapsed coefficients reproduce the reference's two-term update
        collapsed = cx[t] * x_start.double().numpy() - ce[t] * g["p_sample_eps"][i].astype(np.float64) + sg[t] * z.double().numpy()
        assert rel(collapsed, g["p_sample_next"][i]) < 5e-7, t


def test_reverse_coefficients_known_values(golden_dir):
    g = np.load(golden_dir / "ddpm_config.npz")
    cx, ce, sg = O.reverse_coefficients(torch.from_numpy(g["buf_betas"]), torch.from_numpy(g["buf_alphas_cumprod"]))
    assert sg[0] == 0.0
    assert abs(cx[999] - 99.99) < 0.02 and abs(ce[999] - 99.98) < 0.03 and abs(sg[999] - 0.99995) < 1e-4


@pytest.mark.parametrize("name", ["linear3", "smoke"])
def test_full_loop(golden_dir, name):
    g, dims, hidden, schedule, D, seed, sd, x0, cond = load_case(golden_dir, name)
    T = int(g["num_steps"])
    rows = int(g["loop_rows"])
    draw = synth.noise_stream(seed)
    cond_loop = synth.scenario_conditions(rows, dims["condition_dim"]) if dims["condition_dim"] == 3 else cond[:rows]
    final = O.sample(sd, cond_loop, draw(3, (rows, D)), lambda t: draw(10_000 + t, (rows, D)), T)
    # 1000 chained steps: GEMM blocking may differ with the host thread count -> norm-wise 1e-4
    assert rel(final.numpy(), g["loop_final"]) < 1e-4


def test_param_shapes_match_reference_count():
    shapes = synth.param_shapes(5142, 3, (256, 512, 256))
    assert len(shapes) == 52
    assert sum(int(np.prod(s)) for _, s in shapes) == 4_275_798
