"""GPU: the fused bf16 reverse step (fused_step.cuh: output_proj + reverse update + the NEXT step's input_proj in one
kernel, c8 state layout) against the unfused kernels and against the CPU oracle on identical injected noise."""
import pytest
import torch

from oracle import ddpm_oracle as O
from oracle import synth
from tests.helpers import TOL_BF16, build_model, load_case, oracle_sd, rel

pytestmark = pytest.mark.gpu


def _is_fused(model):
    from osteosarcoma_diffusionmodel_b200 import _lib
    return bool(_lib.load().osteo_ddpm_step_is_fused(model._ctx))


@pytest.mark.parametrize("name", ["smoke", "config", "linear3"])
def test_fused_loop_matches_unfused_and_oracle(name):
    case = load_case(name)
    sd = oracle_sd(case)
    T, D = case["T"], case["D"]
    rows, t_stop = 200, T - 12
    draw = synth.noise_stream(4242)
    d = case["dims"]
    cond = synth.make_cohort(rows, d["mutation_dim"], d["expression_dim"], d["pathway_dim"], d["condition_dim"], seed=5)[1]
    x_T = draw(1, (rows, D))
    noises = {t: draw(70_000 + t, (rows, D)) for t in range(t_stop, T)}
    ref = O.sample(sd, cond, x_T, lambda t: noises[t], T, t_stop=t_stop)
    z = torch.stack([noises[t] for t in reversed(range(t_stop, T))])
    model = build_model(case, "bf16")
    model.set_fused(True)
    fused = model.sample(cond, rows, x_T=x_T, noise=z, t_stop=t_stop)
    assert _is_fused(model)
    model.set_fused(False)
    unfused = model.sample(cond, rows, x_T=x_T, noise=z, t_stop=t_stop)
    assert not _is_fused(model)
    model.check_status()
    e_f, e_u, e_fu = rel(fused, ref), rel(unfused, ref), rel(fused, unfused)
    print(f"{name}: fused vs oracle {e_f:.3e}, unfused vs oracle {e_u:.3e}, fused vs unfused {e_fu:.3e}")
    assert e_f < TOL_BF16 and e_u < TOL_BF16
    assert e_fu < TOL_BF16 / 4


def test_fused_single_steps_with_state_reload_and_time_jumps():
    """p_sample reloads the state every call and jumps between timesteps: the priming GEMM has to rerun each time."""
    case = load_case("linear3")
    model = build_model(case, "bf16")
    B, D = case["batch"], case["D"]
    draw = synth.noise_stream(9)
    x = draw(2, (B, D))
    for t in (999, 500, 501, 3, 0):
        z = draw(100 + t, (B, D))
        model.set_fused(True)
        a, ea = model.p_sample(x, t, case["cond"], noise=z, return_eps=True)
        model.set_fused(False)
        b, eb = model.p_sample(x, t, case["cond"], noise=z, return_eps=True)
        assert rel(ea, eb) < 1e-5, t           # eps comes from the same GEMMs in both paths
        assert rel(a, b) < 1e-5, t
    model.check_status()


def test_fused_in_kernel_noise_equals_unfused():
    """Same Philox counters in both paths: the in-kernel-noise loops agree to bf16 tolerance, graph and eager bit-identical."""
    case = load_case("config")
    model = build_model(case, "bf16")
    rows = 300
    cond = synth.scenario_conditions(rows, 3)
    model.set_fused(True)
    model._use_graph = True
    a = model.sample(cond, rows, seed=5, t_stop=990)
    model._use_graph = False
    a2 = model.sample(cond, rows, seed=5, t_stop=990)
    assert torch.equal(a, a2)
    model.set_fused(False)
    b = model.sample(cond, rows, seed=5, t_stop=990)
    assert torch.isfinite(a).all()
    assert rel(a, b) < TOL_BF16 / 4
    model.check_status()


def test_fused_continues_across_sample_loop_calls():
    """Two half loops (the second continues from the internal state through reverse_step) equal one full loop."""
    case = load_case("linear3")
    model = build_model(case, "bf16")
    model.set_fused(True)
    rows = 130
    _, cond = synth.make_cohort(rows, 20, 90, 10, 2, seed=3)
    full = model.sample(cond, rows, seed=3, t_stop=990)
    from osteosarcoma_diffusionmodel_b200 import _lib
    lib, s = _lib.load(), _lib.stream_handle()
    half = model.sample(cond, rows, seed=3, t_stop=995)
    del half
    _lib.check(lib.osteo_ddpm_sample_loop(model._ctx, rows, 994, 990, None, 3, 0, 0, s))
    out = torch.empty((rows, case["D"]), device="cuda", dtype=torch.float32)
    _lib.check(lib.osteo_ddpm_store_state(model._ctx, out.data_ptr(), rows, s))
    assert torch.equal(out, full)
    model.check_status()


@pytest.mark.parametrize("fused", [True, False])
@pytest.mark.parametrize("rows,t", [(300, 500), (129, 1), (64, 999)])
def test_in_kernel_noise_is_the_oracle_philox_stream(fused, rows, t):
    """With every weight zero the denoiser predicts eps = 0, so one reverse step is x <- c_x x + sigma z: z can be read back and must be
    the Philox4x32-10 + Box-Muller normals of the numpy oracle for (seed, global row, column, t) - in both step kernels, with a row
    offset (sharding) and a ragged last row block. Also checks the moments of the extracted noise."""
    import numpy as np
    from oracle import philox_oracle as P
    from osteosarcoma_diffusionmodel_b200.diffusion import BiologyAwareDiffusionModel

    case = load_case("config")
    model = build_model(case, "bf16")
    with torch.no_grad():
        for p_ in model.parameters():
            p_.zero_()
    model.set_fused(fused)
    D = case["D"]
    seed, row_base = 0x1234_5678_9ABC, 1_000_003
    cond = synth.scenario_conditions(rows, 3)
    x = synth.noise_stream(5)(1, (rows, D))
    from osteosarcoma_diffusionmodel_b200 import _lib
    lib, s = _lib.load(), _lib.stream_handle()
    xd = x.cuda().contiguous()
    cd = cond.cuda().contiguous()
    model._ensure_ctx(rows)
    _lib.check(lib.osteo_ddpm_set_conditions(model._ctx, cd.data_ptr(), rows, s))
    _lib.check(lib.osteo_ddpm_load_state(model._ctx, xd.data_ptr(), rows, s))
    _lib.check(lib.osteo_ddpm_reverse_step(model._ctx, rows, t, None, None, seed, row_base, s))
    out = torch.empty_like(xd)
    _lib.check(lib.osteo_ddpm_store_state(model._ctx, out.data_ptr(), rows, s))
    model.check_status()
    cx, ce, sg = BiologyAwareDiffusionModel.reverse_coefficients(model.betas, model.alphas_cumprod)
    z = (out.cpu().double().numpy() - np.float32(cx[t]).astype(np.float64) * x.double().numpy()) / np.float32(sg[t]).astype(np.float64)
    ref = P.normals(seed, np.arange(row_base, row_base + rows, dtype=np.uint64), D, 0, t)      # stream 0 = STREAM_REVERSE
    # fp32 arithmetic of the update on |x| ~ c_x: absolute error ~ 1e-6 * c_x / sigma
    tol = 4e-6 * (1.0 + abs(cx[t])) / sg[t] + 5e-5      # + the MUFU lg2 / sqrt / sin / cos approximations (~1e-6 relative, |z| <= 5.7)
    assert np.abs(z - ref).max() < tol, (np.abs(z - ref).max(), tol)
    assert abs(z.mean()) < 4.0 / np.sqrt(z.size) and abs(z.var() - 1.0) < 6.0 / np.sqrt(z.size)
