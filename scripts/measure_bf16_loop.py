"""Measure (not assert) the error of the benchmarked bf16 mode over the FULL 1000-step loop against the reference's golden
fixtures (injected x_T / z), per checkpoint and for the final sample, plus the thresholded-call agreement. The numbers this prints
are where tests/helpers.py::TOL_BF16_LOOP comes from (profiles/r2_bf16_loop_error.txt)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
from oracle import synth
from tests.helpers import CASES, build_model, load_case, rel

for name in CASES:
    case = load_case(name)
    g = case["g"]
    T, D, rows = case["T"], case["D"], int(g["loop_rows"])
    draw = synth.noise_stream(case["seed"])
    cond = synth.scenario_conditions(rows, 3) if case["dims"]["condition_dim"] == 3 else case["cond"][:rows]
    x_T = draw(3, (rows, D))
    noise = torch.stack([draw(10_000 + t, (rows, D)) if t > 0 else torch.zeros(rows, D) for t in reversed(range(T))])
    for precision in ("bf16", "fp32x3"):
        model = build_model(case, precision)
        line = []
        for ck_t, ck in zip(g["loop_ck_steps"], g["loop_ck"]):
            part = model.sample(cond, rows, x_T=x_T, noise=noise[: T - int(ck_t)], t_stop=int(ck_t))
            line.append(f"t={int(ck_t)}: {rel(part, ck):.3e}")
        final = model.sample(cond, rows, x_T=x_T, noise=noise)
        ref = g["loop_final"]
        md = case["dims"]["mutation_dim"]
        f, r = final.cpu().numpy()[:, :md], ref[:, :md]
        e = rel(final, ref)
        maxabs = float(np.abs(final.cpu().numpy() - ref).max())
        flips = int(((f > 0.5) != (r > 0.5)).sum())
        for tol in (2e-2, 5e-2):
            margin = tol * np.abs(ref).max()
            decided = np.abs(r - 0.5) > margin
            line.append(f"tol {tol}: decided {decided.mean():.3f} flips_in_decided {int(((f > 0.5) != (r > 0.5))[decided].sum())}")
        print(f"{name:8s} {precision:6s} final rel {e:.3e} max|d| {maxabs:.3e} max|ref| {np.abs(ref).max():.3e} flips {flips}/{f.size} | " + " | ".join(line), flush=True)
        model.check_status()
