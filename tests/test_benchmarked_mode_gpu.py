"""GPU: end-to-end parity of the BENCHMARKED mode -- bf16 operands, the fused step kernel, replayed graphs, parallel row branches --
over the full 1000-step loop (models/diffusion.py:427-449): final samples, trajectory checkpoints and thresholded mutation calls
(utils/generate.py:135) against (i) the reference's golden fixtures with injected x_T / z, (ii) rows of a production-size run with the
in-kernel Philox noise against the reference's sample() fed the same streams, (iii) the CPU oracle at config.yaml dims on > 2 row tiles."""
import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as O
from oracle import production_case as PC
from oracle import synth
from tests.helpers import CASES, GOLDEN, TOL_BF16_LOOP, TOL_FP32X3, build_model, load_case, oracle_sd, rel

pytestmark = pytest.mark.gpu


def assert_calls_agree(got: np.ndarray, ref: np.ndarray, tol: float, absmax: float, min_decided: float = 0.8):
    """Thresholded calls must be identical wherever the reference is farther than tol * max|x| from 0.5 (north_star)."""
    decided = np.abs(ref - 0.5) > tol * absmax
    assert decided.mean() > min_decided, decided.mean()
    assert np.array_equal((got > 0.5)[decided], (ref > 0.5)[decided])


@pytest.mark.parametrize("name", list(CASES))
def test_full_loop_bf16_fused_graph_matches_reference_goldens(name):
    """The three reference fixtures (injected x_T and per-step z) through the graph-replayed fused bf16 path: the injected noise stack is
    indexed by the device step word, so this is the same replayed graph the benchmark runs, minus the in-kernel RNG."""
    case = load_case(name)
    g = case["g"]
    T, D, rows = case["T"], case["D"], int(g["loop_rows"])
    draw = synth.noise_stream(case["seed"])
    cond = synth.scenario_conditions(rows, 3) if case["dims"]["condition_dim"] == 3 else case["cond"][:rows]
    x_T = draw(3, (rows, D))
    noise = torch.stack([draw(10_000 + t, (rows, D)) if t > 0 else torch.zeros(rows, D) for t in reversed(range(T))]).cuda()
    model = build_model(case, "bf16")
    assert model._use_graph and model._fused
    for ck_t, ck in zip(g["loop_ck_steps"], g["loop_ck"]):
        part = model.sample(cond, rows, x_T=x_T, noise=noise[: T - int(ck_t)], t_stop=int(ck_t))
        assert rel(part, ck) < TOL_BF16_LOOP, int(ck_t)
    final = model.sample(cond, rows, x_T=x_T, noise=noise)
    mode = model.sampling_mode()
    assert mode == {"precision": "bf16", "fused": True, "graph_branches": 1}, mode
    ref = g["loop_final"]
    assert rel(final, ref) < TOL_BF16_LOOP
    md = case["dims"]["mutation_dim"]
    assert_calls_agree(final.cpu().numpy()[:, :md], ref[:, :md], TOL_BF16_LOOP, float(np.abs(ref).max()))
    # graph replay == eager launches, bit for bit, with injected noise too
    model._use_graph = False
    assert torch.equal(final, model.sample(cond, rows, x_T=x_T, noise=noise))
    model.check_status()


@pytest.mark.parametrize("precision,tol", [("bf16", TOL_BF16_LOOP), ("fp32x3", TOL_FP32X3)])
def test_production_run_matches_the_reference_on_sampled_rows(precision, tol):
    """76 100 patients at config.yaml dims exactly as bench.py samples them -- x_T and z from the in-kernel Philox streams, 10-step and
    1-step replayed graphs, TWO parallel row branches (bf16: the fused kernel) -- against the REFERENCE's own sample() on 208 of those rows
    (first tile, across the branch boundary, ragged last tile; tests/golden/ddpm_production.npz, oracle/gen_golden.py)."""
    g = np.load(GOLDEN / "ddpm_production.npz")
    assert int(g["n_total"]) == PC.N_TOTAL and int(g["seed"]) == PC.SEED and np.array_equal(g["rows"], PC.rows())
    case = load_case("config")
    assert case["seed"] == PC.PARAM_SEED
    model = build_model(case, precision)
    cond = synth.scenario_conditions(PC.N_TOTAL, 3).cuda()
    out = model.sample(cond, PC.N_TOTAL, seed=PC.SEED)
    mode = model.sampling_mode()
    assert mode == {"precision": precision, "fused": precision == "bf16", "graph_branches": 2}, mode
    rows, cols = torch.from_numpy(g["rows"]).cuda(), torch.from_numpy(g["cols"]).cuda()
    got = out[rows][:, cols].cpu().numpy()
    ref = g["final_cols"]
    assert rel(got, ref) < tol
    for part in (slice(0, 64), slice(64, 144), slice(144, 208)):            # every row group on its own: no tile may hide behind the others
        assert rel(got[part], ref[part]) < tol
    md = case["dims"]["mutation_dim"]
    assert_calls_agree(got[:, :md], ref[:, :md], tol, float(g["final_absmax"]))
    # the same rows sampled ALONE (one tile-sized call, single branch) are bit-identical: rows do not depend on batch composition
    sub = slice(64, 144)
    alone = model.sample(cond[rows[sub]], 80, seed=PC.SEED, row_base=int(g["rows"][64]))
    assert torch.equal(alone[:, cols], out[rows[sub]][:, cols])
    model.check_status()


def test_bf16_graph_loop_matches_oracle_at_config_dims_on_three_row_tiles():
    """300 rows (two full 128-row tiles + a ragged one) at config.yaml dims, full loop, injected noise: benchmarked path vs the CPU oracle."""
    case = load_case("config")
    sd = oracle_sd(case)
    T, D, rows = case["T"], case["D"], 300
    gen = torch.Generator().manual_seed(4242)
    cond = synth.scenario_conditions(rows, 3)
    x_T = torch.randn((rows, D), generator=gen)
    noise = torch.randn((T, rows, D), generator=gen)
    noise[T - 1].zero_()                                     # t = 0 draws nothing (models/diffusion.py:408)
    ref = O.sample(sd, cond, x_T, lambda t: noise[T - 1 - t], T)
    model = build_model(case, "bf16")
    got = model.sample(cond, rows, x_T=x_T, noise=noise.cuda())
    assert model.sampling_mode()["fused"] and model.sampling_mode()["graph_branches"] == 1
    assert rel(got, ref) < TOL_BF16_LOOP
    for r0 in (0, 128, 256):
        assert rel(got[r0:r0 + 128], ref[r0:r0 + 128]) < TOL_BF16_LOOP, r0
    md = case["dims"]["mutation_dim"]
    assert_calls_agree(got.cpu().numpy()[:, :md], ref.numpy()[:, :md], TOL_BF16_LOOP, float(ref.abs().max()))
    model.check_status()
