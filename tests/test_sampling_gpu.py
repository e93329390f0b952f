"""GPU: the reverse process (p_sample / sample) and the denoiser forward against golden fixtures written by the
reference itself and against the CPU oracle on the same injected noise."""
import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as O
from oracle import synth
from tests.helpers import CASES, TOL_BF16, TOL_FP32X3, build_model, load_case, oracle_sd, rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("precision,tol", [("fp32x3", TOL_FP32X3), ("bf16", TOL_BF16)])
def test_single_reverse_steps_match_reference(name, precision, tol):
    case = load_case(name)
    g = case["g"]
    model = build_model(case, precision)
    draw = synth.noise_stream(case["seed"])
    B, D = case["batch"], case["D"]
    x_start = draw(2, (B, D)) * 1.5
    for i, t in enumerate(g["p_sample_steps"]):
        t = int(t)
        z = draw(100 + t, (B, D))
        nxt, eps = model.p_sample(x_start, t, case["cond"], noise=z if t > 0 else None, return_eps=True)
        assert rel(eps, g["p_sample_eps"][i]) < tol, (t, "eps")
        # the update multiplies eps by c_eps up to ~100 at t=999 while x_{t-1} stays O(|x| + |eps|): same norm-wise bound
        assert rel(nxt, g["p_sample_next"][i]) < tol, (t, "x_next")
    model.check_status()


@pytest.mark.parametrize("name", list(CASES))
def test_denoiser_forward_matches_reference(name):
    case = load_case(name)
    g = case["g"]
    model = build_model(case, "fp32x3")
    noise = synth.noise_stream(case["seed"])(1, (case["batch"], case["D"]))
    t = torch.from_numpy(g["t_idx"])
    x_t, n2 = model.q_sample(case["x0"].cuda(), t.cuda(), noise.cuda())
    assert np.array_equal(x_t.cpu().numpy(), g["q_sample"])            # elementwise path: bit-exact
    assert torch.equal(n2.cpu(), noise)
    model._inject = {"t": t, "noise": noise}
    eps = model(case["x0"].cuda(), case["cond"].cuda(), return_loss=False)
    assert rel(eps, g["eps_hat"]) < TOL_FP32X3
    eps2 = model.predict_noise(x_t, t, case["cond"])
    assert torch.equal(eps, eps2)
    model.set_precision("bf16")
    assert rel(model.predict_noise(x_t, t, case["cond"]), g["eps_hat"]) < TOL_BF16
    model.check_status()


@pytest.mark.parametrize("name", list(CASES))
def test_full_1000_step_loop_matches_reference(name):
    case = load_case(name)
    g = case["g"]
    T, D, rows = case["T"], case["D"], int(g["loop_rows"])
    draw = synth.noise_stream(case["seed"])
    cond = synth.scenario_conditions(rows, 3) if case["dims"]["condition_dim"] == 3 else case["cond"][:rows]
    x_T = draw(3, (rows, D))
    noise = torch.stack([draw(10_000 + t, (rows, D)) if t > 0 else torch.zeros(rows, D) for t in reversed(range(T))])
    model = build_model(case, "fp32x3")
    # checkpoints of the trajectory, then the final sample
    for ck_t, ck in zip(g["loop_ck_steps"], g["loop_ck"]):
        ck_t = int(ck_t)
        part = model.sample(cond, rows, x_T=x_T, noise=noise[: T - ck_t], t_stop=ck_t)
        assert rel(part, ck) < TOL_FP32X3, ck_t
    final = model.sample(cond, rows, x_T=x_T, noise=noise)
    ref = g["loop_final"]
    assert rel(final, ref) < TOL_FP32X3
    # thresholded mutation calls (utils/generate.py:135) agree wherever the reference is not within tolerance of 0.5
    md = case["dims"]["mutation_dim"]
    f, r = final.cpu().numpy()[:, :md], ref[:, :md]
    margin = TOL_FP32X3 * np.abs(ref).max()
    decided = np.abs(r - 0.5) > margin
    assert decided.mean() > 0.9
    assert np.array_equal((f > 0.5)[decided], (r > 0.5)[decided])
    model.check_status()


def test_loop_matches_cpu_oracle_on_fresh_inputs():
    """Oracle and CUDA path on the same seeded inputs that are NOT in the fixtures (different seed / batch)."""
    case = load_case("linear3")
    sd = oracle_sd(case)
    T, D = case["T"], case["D"]
    rows = 7
    draw = synth.noise_stream(991)
    _, cond = synth.make_cohort(rows, 20, 90, 10, 2, seed=55)
    x_T = draw(1, (rows, D))
    t_stop = 900
    noises = {t: draw(50_000 + t, (rows, D)) for t in range(t_stop, T)}
    ref = O.sample(sd, cond, x_T, lambda t: noises[t], T, t_stop=t_stop)
    model = build_model(case, "fp32x3")
    z = torch.stack([noises[t] for t in reversed(range(t_stop, T))])
    got = model.sample(cond, rows, x_T=x_T, noise=z, t_stop=t_stop)
    assert rel(got, ref) < TOL_FP32X3
    model.set_precision("bf16")
    got_bf = model.sample(cond, rows, x_T=x_T, noise=z, t_stop=t_stop)
    assert rel(got_bf, ref) < TOL_BF16


@pytest.mark.parametrize("rows", [1, 127, 129, 300])
def test_ragged_batches_and_graph_equals_eager(rows):
    case = load_case("linear3")
    model = build_model(case, "bf16")
    _, cond = synth.make_cohort(rows, 20, 90, 10, 2, seed=3)
    model._use_graph = True
    a = model.sample(cond, rows, seed=11, t_stop=980)
    model._use_graph = False
    b = model.sample(cond, rows, seed=11, t_stop=980)
    assert torch.equal(a, b)
    assert torch.isfinite(a).all()
    model.check_status()


def test_rows_are_independent_of_sharding_and_chunking():
    """Philox is keyed by the GLOBAL row: any split of the cohort over GPUs / chunks yields identical rows (§8e)."""
    case = load_case("linear3")
    model = build_model(case, "bf16")
    n = 384
    _, cond = synth.make_cohort(n, 20, 90, 10, 2, seed=4)
    full = model.sample(cond, n, seed=5, t_stop=990)
    lo = model.sample(cond[:130], 130, seed=5, row_base=0, t_stop=990)
    hi = model.sample(cond[130:], n - 130, seed=5, row_base=130, t_stop=990)
    assert torch.equal(full, torch.cat([lo, hi]))
    model.set_chunk_rows(128)
    assert torch.equal(full, model.sample(cond, n, seed=5, t_stop=990))
    other = model.sample(cond, n, seed=6, t_stop=990)
    assert not torch.equal(full, other)


def test_in_kernel_noise_statistics():
    """With the denoiser's output_proj zeroed, one step from x=0 is sigma[t] * z: checks the fused Philox epilogue."""
    case = load_case("linear3")
    sd = dict(case["sd"])
    sd["unet.output_proj.weight"] = torch.zeros_like(sd["unet.output_proj.weight"])
    sd["unet.output_proj.bias"] = torch.zeros_like(sd["unet.output_proj.bias"])
    case = dict(case, sd=sd)
    model = build_model(case, "bf16")
    n, D, t = 4096, case["D"], 500
    _, cond = synth.make_cohort(n, 20, 90, 10, 2, seed=4)
    out = model.p_sample(torch.zeros(n, D), t, cond, seed=77)
    cx, ce, sg = O.reverse_coefficients(model.betas.cpu(), model.alphas_cumprod.cpu())
    z = out / sg[t]
    assert abs(z.mean().item()) < 5e-3 and abs(z.var().item() - 1.0) < 5e-3
    from oracle import philox_oracle as P
    ref = P.normals(77, np.arange(8, dtype=np.uint64), D, 0, t) * np.float32(sg[t])
    assert np.abs(out[:8].cpu().numpy() - ref).max() < 1e-5


def test_standalone_reverse_update_matches_fused_epilogue():
    import ctypes as C
    from osteosarcoma_diffusionmodel_b200 import _lib

    case = load_case("smoke")
    model = build_model(case, "fp32x3")
    draw = synth.noise_stream(5)
    B, D, t = 4, case["D"], 700
    x, z = draw(1, (B, D)).cuda(), draw(2, (B, D)).cuda()
    nxt, eps = model.p_sample(x, t, case["cond"], noise=z, return_eps=True)
    x2 = x.clone()
    _lib.check(_lib.load().osteo_ddpm_reverse_update(model._ctx, x2.data_ptr(), eps.data_ptr(), z.data_ptr(), B, t, 0, 0, _lib.stream_handle()))
    assert torch.equal(x2, nxt)


@pytest.mark.parametrize("hidden", [(128, 128), (512, 256, 128, 256), (256, 256, 256)])
def test_other_topologies_match_the_oracle(hidden):
    """The constructor is generic in hidden_dims (models/diffusion.py:171-193): encoder / bottleneck / decoder wiring and the
    GroupNorm group widths 16 / 32 / 64 against the oracle for a short injected-noise loop and one training step."""
    dims = dict(mutation_dim=12, expression_dim=100, pathway_dim=8, condition_dim=3)
    D, T, rows = 120, 1000, 70
    sd = synth.make_params(D, 3, hidden, seed=11)
    osd = dict(sd)
    osd.update(O.schedule_buffers("cosine", T))
    from osteosarcoma_diffusionmodel_b200.diffusion import BiologyAwareDiffusionModel
    model = BiologyAwareDiffusionModel(12, 100, 8, 3, synth.model_config(hidden_dims=hidden))
    model.load_state_dict(sd, strict=False)
    model = model.cuda().eval().set_precision("fp32x3")
    draw = synth.noise_stream(17)
    cond = synth.scenario_conditions(rows, 3)
    x_T = draw(1, (rows, D))
    t_stop = 990
    noises = {t: draw(100 + t, (rows, D)) for t in range(t_stop, T)}
    ref = O.sample(osd, cond, x_T, lambda t: noises[t], T, t_stop=t_stop)
    got = model.sample(cond, rows, x_T=x_T, noise=torch.stack([noises[t] for t in reversed(range(t_stop, T))]), t_stop=t_stop)
    assert rel(got, ref) < TOL_FP32X3
    # one training step: loss and every gradient
    x0, c = synth.make_cohort(rows, 12, 100, 8, 3, seed=2)
    t = torch.from_numpy(np.random.RandomState(3).randint(0, T, size=rows).astype(np.int64))
    noise = draw(7, (rows, D))
    masks = synth.dropout_masks(5, rows, synth.block_widths(hidden), 0.2)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    full = dict(osd)
    full.update(params)
    ref_loss = O.forward_loss(full, x0, c, t, noise, T, drop_masks=masks, p=0.2, training=True)
    ref_loss.backward()
    model.train()
    model._inject = {"t": t, "noise": noise, "masks": masks}
    loss = model(x0.cuda(), c.cuda())
    loss.backward()
    assert abs(loss.item() - ref_loss.item()) < 1e-5 * abs(ref_loss.item())
    for n, p in model.named_parameters():
        assert rel(p.grad, params[n].grad) < 2e-4, n
    model.check_status()


def test_more_rows_than_one_chunk_and_single_row():
    case = load_case("linear3")
    model = build_model(case, "bf16")
    model.set_chunk_rows(256)
    n = 1000
    _, cond = synth.make_cohort(n, 20, 90, 10, 2, seed=8)
    a = model.sample(cond, n, seed=3, t_stop=995)
    model.set_chunk_rows(0)
    b = model.sample(cond, n, seed=3, t_stop=995)
    assert torch.equal(a, b)
    one = model.sample(cond[:1], 1, seed=3, t_stop=995)
    assert torch.equal(one, a[:1])
    with pytest.raises(ValueError):
        model.sample(cond[:3], 5)
    model.check_status()


def test_sample_components_is_the_split_and_thresholded_sample():
    """utils/generate.py:130-135 on the device: mutation calls bit-identical to (sample()[:, :mutation_dim] > 0.5), bit packing LSB
    first, expression / pathway blocks equal to the column split."""
    case = load_case("smoke")
    model = build_model(case, "bf16")
    n = 300
    cond = case["cond"][:1].repeat(n, 1).cuda()
    full = model.sample(cond, n, seed=5, t_stop=900)
    comp = model.sample_components(cond, n, seed=5, t_stop=900, pack_bits=True)
    md, ed = model.mutation_dim, model.expression_dim
    calls = (full[:, :md] > 0.5)
    assert comp["mutations"].dtype == torch.uint8 and torch.equal(comp["mutations"].bool(), calls)
    assert torch.equal(comp["expression"], full[:, md:md + ed]) and torch.equal(comp["pathways"], full[:, md + ed:])
    bits = comp["mutation_bits"].cpu().numpy()
    unpacked = np.unpackbits(bits, axis=1, bitorder="little")[:, :md]
    assert np.array_equal(unpacked.astype(bool), calls.cpu().numpy())
    assert 0 < calls.float().mean().item() < 1       # both outcomes occur
    model.check_status()


def test_parallel_row_branches_do_not_change_the_samples():
    """Sampling graphs split big batches into parallel row branches with their own step words (osteo_ddpm_set_branches; two waves of row
    tiles per branch are required, i.e. >= 75 776 rows on 148 SMs): bit-identical to the single-branch graph, ragged last tile included,
    across a 10-step unrolled graph plus single-step graphs."""
    case = load_case("config")
    model = build_model(case, "bf16")
    n = 76_001
    cond = synth.scenario_conditions(n, 3).cuda()
    outs = []
    for nb in (1, 2, 1):
        model.set_branches(nb)
        outs.append(model.sample(cond, n, seed=3, t_stop=987)[::97].clone())       # 13 steps = one 10-step graph + three 1-step graphs
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert torch.isfinite(outs[0]).all()
    model.check_status()


def test_shard_writer_matches_sample_components(tmp_path):
    """egress.generate_to_shards (pinned double buffer + writer thread) on the device path: the files of a 3-shard run equal one
    sample_components call over the whole cohort (rows keep their global Philox identity), bits and 0/1 bytes alike."""
    from osteosarcoma_diffusionmodel_b200.egress import generate_to_shards, load_shards

    case = load_case("smoke")
    model = build_model(case, "bf16")
    n = 333
    cond = case["cond"][:1].repeat(n, 1).cuda() + torch.arange(n, device="cuda")[:, None] * 1e-3
    whole = model.sample_components(cond, n, seed=8, row_base=50, t_stop=950)
    for shard_rows, bits in [(128, True), (200, False)]:
        d = tmp_path / f"s{shard_rows}"
        # t_stop is a test-only shortcut: drive the writer through a thin wrapper that fixes it
        class Short:
            mutation_dim, expression_dim, pathway_dim = model.mutation_dim, model.expression_dim, model.pathway_dim
            def sample_components(self, c, m, **kw):
                return model.sample_components(c, m, t_stop=950, **kw)
        man = generate_to_shards(Short(), cond, d, shard_rows=shard_rows, seed=8, row_base=50, pack_bits=bits)
        assert len(man["shards"]) == -(-n // shard_rows)
        got = load_shards(d)
        assert np.array_equal(got["mutations"], whole["mutations"].cpu().numpy().astype(float))
        assert np.array_equal(got["expression"], whole["expression"].cpu().numpy())
        assert np.array_equal(got["pathways"], whole["pathways"].cpu().numpy())
        assert np.array_equal(got["conditions"], cond.cpu().numpy())
    model.check_status()


def test_block_gemm_kernel_variants_agree(monkeypatch):
    """The Linear+GroupNorm+SiLU layers of the bf16 mode run on three kernels -- the generic streaming kernel, the weight-stationary kernel
    (lean epilogue, per-warp TMA stores for K <= 384, st.global otherwise) and the CTA-pair (cta_group::2) kernel for the 512 x 512 layers.
    Same inputs through every combination, at a row count that is neither a whole number of 128-row blocks nor an even number of them (the
    last pair of the CTA-pair kernel has one live block): they may differ by bf16 rounding of the activations only, and each stays within
    the single-call bf16 tolerance of the fp32x3 result."""
    case = load_case("config")
    n = 5 * 128 - 17
    rs = np.random.RandomState(11)
    x_t = torch.from_numpy(rs.standard_normal((n, case["D"])).astype(np.float32)).cuda()
    t = torch.from_numpy(rs.randint(0, case["T"], size=n).astype(np.int64)).cuda()
    cond = synth.scenario_conditions(n, 3).cuda()
    ref = build_model(case, "fp32x3").predict_noise(x_t, t, cond)
    outs = {}
    for name, env in {"default": {}, "one_cta_only": {"OSTEO_WS2": "0"}, "pairs_everywhere": {"OSTEO_WS2": "2"}, "st_global": {"OSTEO_WS_TMA_STORE": "0"},
                      "generic": {"OSTEO_DDPM_NO_WS": "1"}}.items():
        for k in ("OSTEO_WS2", "OSTEO_WS_TMA_STORE", "OSTEO_DDPM_NO_WS"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        model = build_model(case, "bf16")          # the switches are read when the context is created
        outs[name] = model.predict_noise(x_t, t, cond)
        model.check_status()
        assert rel(outs[name], ref) < TOL_BF16, name
    for name, o in outs.items():
        assert rel(o, outs["generic"]) < 5e-3, name
