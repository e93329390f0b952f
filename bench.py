#!/usr/bin/env python
"""Benchmark of the B200-native conditional-DDPM sampling path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--rows R]

Metric (BASELINE.json): DDPM-sampled patients/sec, 1000 reverse steps per patient.
Workload at N=1 (BASELINE.json configs[1]): 1000-step conditional DDPM sampling of 100k synthetic
patients, bf16 tensor-core operands, the three config.yaml scenarios, config.yaml dims
(62 + 5054 + 26 = 5142 features, 3 conditions, hidden [256, 512, 256], cosine schedule).
One bench "step" = one complete `model.sample()` of the whole per-GPU batch (1000 reverse steps).
N > 1 (torchrun, one rank per GPU): every rank samples its own contiguous row range of the global
cohort (Philox keyed by global row; no collective on the data path) -> weak scaling.

`--impl reference` times the reference's CPU algorithm (the torch-CPU oracle port of
models/diffusion.py:382-449, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

D_MUT, D_EXPR, D_PATH, N_COND = 62, 5054, 26, 3            # config/config.yaml:27-30
D = D_MUT + D_EXPR + D_PATH
T_STEPS = 1000
HIDDEN = (256, 512, 256)
ALGO_BYTES_PER_PATIENT_STEP = 2 * D * 4                     # read x_t once, write x_{t-1} once, fp32 state
ALGO_FLOPS_PER_PATIENT_STEP = 2 * 4_205_568                 # 12 core GEMMs
METRIC = "ddpm_sampled_patients_per_sec_1000_steps"
UNIT = "patients/s"


def read_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1400.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.lines = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu_index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.lines.append(line.strip())

        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power) if power else None}


# ----------------------------------------------------------------------------- CPU reference arm
def reference_model(device="cpu"):
    """The UNMODIFIED reference class (models/diffusion.py:259-449) from the staged copy oracle/_ref/reference (oracle/stage_reference.py;
    /root/reference in the build container), random-init weights of the bench workload. None when no copy is present."""
    import torch
    from oracle import reference_import as R
    from osteosarcoma_diffusionmodel_b200 import synthetic as synth

    if not R.available():
        return None
    ref_diffusion, _ = R.import_reference()
    m = ref_diffusion.BiologyAwareDiffusionModel(D_MUT, D_EXPR, D_PATH, N_COND, synth.model_config(hidden_dims=HIDDEN))
    m.load_state_dict(synth.make_params(D, N_COND, HIDDEN, seed=0), strict=False)
    return m.to(device).eval()


def cpu_reference_rate(rows: int, steps: int, repeats: int = 1, model=None):
    """patients/s of the reference's CPU path: `steps` reverse steps (p_sample, models/diffusion.py:382-425; noise drawn inside the timed
    region as the reference does) on `rows` rows, torch CPU with all host threads, scaled to the 1000 steps a patient needs.
    Runs the staged reference itself (kind "reference") or, without it, the oracle port (kind "port")."""
    import torch
    from osteosarcoma_diffusionmodel_b200 import synthetic as synth

    torch.set_num_threads(os.cpu_count() or 1)
    cond = synth.scenario_conditions(rows, N_COND)
    x = torch.randn(rows, D)
    if model is None:
        model = reference_model()
    if model is not None:
        kind = "reference"

        def one(xx, t):
            return model.p_sample(xx, t, cond)
    else:
        from oracle import ddpm_oracle as O

        kind = "port"
        sd = synth.make_params(D, N_COND, HIDDEN, seed=0)
        sd.update(O.schedule_buffers("cosine", T_STEPS))

        def one(xx, t):
            return O.p_sample(sd, xx, t, cond, torch.randn_like(xx), T_STEPS)
    best = None
    for _ in range(repeats):
        xx = x.clone()
        t0 = time.perf_counter()
        for i in range(steps):
            xx = one(xx, T_STEPS - 1 - i)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    per_step = best / steps
    return rows / (per_step * T_STEPS), per_step, torch.get_num_threads(), kind


def run_reference_arm(args):
    """The reference's own CPU implementation of the path on this box's host cores (all threads): every bench step is a bounded sample
    -- 1024 rows (the reference's CPU throughput peak) x 20 of the 1000 reverse steps, scaled -- plus, once, a TRUE
    1000-step sample() of 100 patients (BASELINE.json configs[0]'s sampling half) and compute_mmd at N = 2000."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import torch
    from osteosarcoma_diffusionmodel_b200 import synthetic as synth

    rows, sub_steps = 1024, 20
    model = reference_model()
    vals, ms = [], []
    for _ in range(args.warmup):
        cpu_reference_rate(rows, 2, model=model)
    for _ in range(args.steps):
        v, per_step, threads, kind = cpu_reference_rate(rows, sub_steps, model=model)
        vals.append(v)
        ms.append(per_step * sub_steps * 1e3)
    value = statistics.mean(vals)
    sample = (f"{rows} rows x {sub_steps} of {T_STEPS} reverse steps per bench step, scaled to {T_STEPS} steps (B=1024 is the reference's CPU "
              f"throughput peak, BASELINE.md §2); {'the staged reference itself (oracle/_ref/reference)' if kind == 'reference' else 'oracle port (no staged reference found)'}")
    secondary = {}
    if model is not None and not args.no_secondary:
        cond = synth.scenario_conditions(100, N_COND)
        t0 = time.perf_counter()
        with torch.no_grad():
            out = model.sample(cond, num_samples=100)
        dt = time.perf_counter() - t0
        secondary["true_1000_step_sample"] = {"rows": 100, "s": dt, "patients_per_s": 100 / dt, "finite": bool(torch.isfinite(out).all()),
                                             "note": "model.sample(conditions, 100): the full 1000-step loop, not scaled"}
        try:
            import numpy as np
            from oracle import reference_import as R

            _, ref_validation = R.import_reference()
            val = ref_validation.BiologicalValidator({"evaluation": {"driver_genes": [], "mutually_exclusive_pairs": [], "required_correlations": []}})
            rs = np.random.RandomState(0)
            n = 2000
            X, Y = rs.standard_normal((n, D)) + 4.0, rs.standard_normal((n, D)) * 1.1 + 4.1
            t0 = time.perf_counter()
            mmd = val.compute_mmd(X, Y)
            dt = time.perf_counter() - t0
            secondary["compute_mmd"] = {"rows": n, "features": D, "s": dt, "kernel_pairs_per_s": 3.0 * n * n / dt, "value": float(mmd),
                                        "note": "utils/validation.py:273-298 (scipy cdist, float64); O(N^2): 1 M x 1 M extrapolates to ~%.0f days" % (3.0e12 / (3.0 * n * n / dt) / 86400)}
        except Exception as e:
            secondary["compute_mmd"] = {"error": repr(e)}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": statistics.mean(ms), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "1000-step conditional DDPM sampling, config.yaml dims (5142 features, 3 conditions, hidden [256,512,256], cosine), three config.yaml "
                               "scenarios (BASELINE.json configs[1]) -- CPU arm: a bounded sample of it per bench step",
                   "rows_per_bench_step": rows, "reverse_steps_per_bench_step": sub_steps, "num_steps": T_STEPS, "precision": "fp32 (torch CPU eager)",
                   "rng": "torch.randn_like inside p_sample", "weights": "random init (synthetic.make_params seed 0)",
                   "value_is": "rows / (seconds per reverse step x 1000)", "ms_per_step_is": f"wall time of the {sub_steps} reverse steps actually run"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "secondary": secondary,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(args, rows_per_gpu):
    return {"workload": "1000-step conditional DDPM sampling, config.yaml dims (5142 features, 3 conditions, hidden [256,512,256], cosine), "
                        "three config.yaml scenarios in equal thirds (BASELINE.json configs[1])",
            "patients_per_gpu": rows_per_gpu, "num_steps": T_STEPS, "precision": args.precision,
            "rng": "in-kernel Philox4x32-10 keyed by (seed, global row, t, column)", "weights": "random init (synthetic.make_params seed 0)",
            "l2_policy": "fp32 state (2.1 GB per 100k patients) is larger than L2; no flush needed",
            "sharding": "contiguous global-row ranges per rank, no data-path collective", "graph": "10-step executable graphs, 2 parallel row branches per rank (osteo_ddpm_set_branches)"}


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device. The product path has no CPU fallback; use --impl reference for the CPU arm.")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        # NCCL prints its version banner on stdout at communicator creation: keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    from osteosarcoma_diffusionmodel_b200 import build
    from osteosarcoma_diffusionmodel_b200.diffusion import BiologyAwareDiffusionModel
    from osteosarcoma_diffusionmodel_b200 import _lib
    from osteosarcoma_diffusionmodel_b200 import synthetic as synth   # deterministic synthetic weights / cohort generator

    build.build()
    rows = args.rows
    cfg = synth.model_config(hidden_dims=HIDDEN)
    model = BiologyAwareDiffusionModel(D_MUT, D_EXPR, D_PATH, N_COND, cfg)
    model.load_state_dict(synth.make_params(D, N_COND, HIDDEN, seed=0), strict=False)
    model = model.to(dev).eval()
    model.set_precision(args.precision)
    if args.chunk_rows is not None:
        model.set_chunk_rows(args.chunk_rows)
    row_base = rank * rows
    cond_host = synth.scenario_conditions(rows, N_COND).pin_memory()
    cond_dev = cond_host.to(dev)
    out_host = torch.empty((rows, D), dtype=torch.float32).pin_memory() if not args.no_e2e else None

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(seed):
        return model.sample(cond_dev, rows, seed=seed, row_base=row_base)

    for i in range(args.warmup):
        one_step(1000 + i)
    model.check_status()

    # ---- kernel-resident timing: inputs already in HBM, result left in HBM
    sampler = ClockSampler(local_rank)
    launches0 = model.launch_count()
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        one_step(i)
    ev1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    barrier()
    launches = model.launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    if dist is not None:
        tt = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = tt.item()
    model.check_status()
    value = rows * world * args.steps / (ms_total / 1e3)

    # ---- end to end through the public API: pinned host conditions in, samples out to pinned host memory
    e2e = None
    if not args.no_e2e:
        def e2e_step(seed):
            c = cond_host.to(dev, non_blocking=True)
            s = model.sample(c, rows, seed=seed, row_base=row_base)
            out_host.copy_(s, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        e2e_step(77)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            e2e_step(100 + i)
        e1.record()
        torch.cuda.synchronize()
        ms_e2e = e0.elapsed_time(e1)
        if dist is not None:
            tt = torch.tensor([ms_e2e], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms_e2e = tt.item()
        e2e = {"value": rows * world * args.steps / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(cond_host.numel() * 4),
               "d2h_bytes_per_step": int(rows * D * 4)}

    # ---- live per-kernel timing of one reverse step (CUDA events after every launch on the launch stream)
    roofline = None
    kernels = None
    hbm_peak, tf_peak, peak_src = read_peaks()
    if rank == 0:
        import ctypes as C

        lib = _lib.load()
        buf = (C.c_float * 4096)()
        per = []
        # the clocks of a power-capped board depend on what ran just before: bring them to the loaded state of the timed region first
        model.sample(cond_dev, rows, seed=5, row_base=row_base, t_stop=T_STEPS - 100)
        for rep in range(10):
            n_l = lib.osteo_ddpm_profile_step(model._ctx, rows, 500, 9, row_base, buf, 4096, _lib.stream_handle())
            if n_l < 0:
                _lib.check(n_l)
            per.append([buf[i] for i in range(n_l)])
        per = np.median(np.array(per[2:]), axis=0)      # drop the first repetitions, median of the rest
        fused = bool(lib.osteo_ddpm_step_is_fused(model._ctx))
        if fused:      # fused_step.cuh: output_proj + reverse update + the next step's input_proj are one kernel
            names = [f"linear_gn_silu_{i}" for i in range(10)] + ["output_proj+reverse_update+next_input_proj"]
        else:
            names = ["input_proj+emb_add"] + [f"linear_gn_silu_{i}" for i in range(10)] + ["output_proj+reverse_update"]
        by_kernel = {}
        for i, ms in enumerate(per):
            by_kernel.setdefault(names[i % len(names)], []).append(float(ms))
        step_ms = float(per.sum())
        kernels = {k: {"ms_per_step": sum(v), "share": sum(v) / step_ms} for k, v in by_kernel.items()}
        dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
        launches_of_dom = len(by_kernel[dom])
        dom_ms = kernels[dom]["ms_per_step"] / launches_of_dom
        rows_per_launch = rows / launches_of_dom
        if dom.startswith("output_proj+reverse_update"):
            algo = ALGO_BYTES_PER_PATIENT_STEP * rows_per_launch
            ach = algo / (dom_ms / 1e3) / 1e9
            roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                        "algorithmic_bytes_per_launch": algo, "avg_launch_ms": dom_ms, "peak_source": peak_src,
                        "timing": "avg_launch_ms = (per-step time of the timed region, CUDA events) x (the kernel's share of one eagerly launched reverse step "
                                  "over all rows, CUDA events around every launch: osteo_ddpm_profile_step, median of 8 repetitions after a 100-step warm-up); "
                                  "inside the timed region the kernels are graph nodes, two row branches side by side; 'standalone' = the eager launch itself"}
            # Headline figure = the kernel's time INSIDE THE TIMED REGION: per-step time of the replayed graphs x the kernel's share of the
            # eagerly launched step (events cannot be recorded between the nodes of a replayed graph). The eager per-launch time itself
            # swings by +-6 % with the clock state a power-capped board happens to be in when the short probe runs; it is kept under
            # "standalone".
            loop_step_ms = ms_total / args.steps / T_STEPS
            in_loop_ms = kernels[dom]["share"] * loop_step_ms / launches_of_dom
            roofline["standalone"] = {"avg_launch_ms": dom_ms, "achieved": ach, "frac": ach / hbm_peak, "ms_per_reverse_step": step_ms}
            roofline["avg_launch_ms"] = in_loop_ms
            roofline["achieved"] = algo / (in_loop_ms / 1e3) / 1e9
            roofline["frac"] = roofline["achieved"] / hbm_peak
            roofline["in_timed_region"] = {"ms_per_reverse_step": loop_step_ms, "kernel_share_of_step": kernels[dom]["share"]}
        elif dom == "input_proj+emb_add":
            algo = D * 4 * rows_per_launch      # one read of the state row (fp32-equivalent algorithmic bytes)
            ach = algo / (dom_ms / 1e3) / 1e9
            roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": None,
                        "algorithmic_bytes_per_launch": algo, "avg_launch_ms": dom_ms, "peak_source": peak_src}
        else:
            flops = ALGO_FLOPS_PER_PATIENT_STEP * rows_per_launch / 12      # coarse: one of 12 equal shares
            ach = flops / (dom_ms / 1e3) / 1e12
            roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s", "frac": ach / tf_peak, "traffic": None,
                        "avg_launch_ms": dom_ms, "peak_source": peak_src}
        prof = ROOT / "profiles" / "traffic.json"
        if prof.exists():
            try:
                roofline["traffic"] = json.loads(prof.read_text()).get(dom)
            except Exception:
                pass
        # whole reverse step against both ceilings of SURVEY.md §8(d)
        gemm_ms = sum(v["ms_per_step"] for k, v in kernels.items())
        roofline["step"] = {"ms_per_reverse_step": step_ms, "hbm_frac_of_algorithmic": ALGO_BYTES_PER_PATIENT_STEP * rows / (step_ms / 1e3) / 1e9 / hbm_peak,
                            "tensor_frac": ALGO_FLOPS_PER_PATIENT_STEP * rows / (gemm_ms / 1e3) / 1e12 / tf_peak}

    # ---- secondary workloads of the same hot path (BASELINE.json: "train step time"; configs[3], configs[4]); N=1 only
    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary:
        secondary = run_secondary(model, dev, hbm_peak, tf_peak)

    if world > 1 and not args.no_secondary:
        secondary = run_secondary_dp(model, dev, dist, rank, world, tf_peak)      # every rank takes part; rank 0 keeps the numbers

    # ---- CPU baseline on this box's host cores (rank 0, N=1 only; bounded sample)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, per_step, threads, kind = cpu_reference_rate(1024, 20)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": kind,
                        "sample": f"1024 rows x 20 of {T_STEPS} reverse steps of " + ("the staged reference's own p_sample (oracle/_ref/reference)" if kind == "reference"
                                  else "the torch-CPU oracle port (oracle/ddpm_oracle.py)") + f", scaled to {T_STEPS} steps"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "bf16x3 (split-bf16, fp32-equivalent)", "data": "synthetic",
            "config": workload_config(args, rows), "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu_baseline, "secondary": secondary,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_secondary(model, dev, hbm_peak, tf_peak):
    """Train step (fwd + bwd + clip + AdamW, batch 8192 = one GPU's share of BASELINE.json configs[3]), RBF-MMD Gram reduction and
    pathway-coherence moments, each timed with CUDA events after warm-up. CPU figures are the oracle port on bounded samples."""
    import numpy as np
    import torch
    from oracle import ddpm_oracle as O            # CPU-port legs only
    from oracle import validators_oracle as V      # CPU-port legs only
    from osteosarcoma_diffusionmodel_b200 import _lib
    from osteosarcoma_diffusionmodel_b200 import synthetic as synth
    from osteosarcoma_diffusionmodel_b200.validation import BiologicalValidator

    out = {}

    def timed(fn, warm, reps):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # -------- training step
    B = 8192
    x0, cond = synth.make_cohort(B, D_MUT, D_EXPR, D_PATH, N_COND, seed=3)
    x0, cond = x0.to(dev), cond.to(dev)
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)      # utils/train.py:169-173

    def step():
        opt.zero_grad()                       # utils/train.py:230
        loss = model(x0, cond, return_loss=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()

    def fwd_bwd():
        model.zero_grad()
        model(x0, cond, return_loss=True).backward()

    ms = timed(step, 3, 10)
    ms_fb = timed(fwd_bwd, 2, 10)
    flops = 22.6e6 * B          # fwd + wgrad + dgrad per sample
    out["train_step"] = {"batch": B, "precision": model._precision, "ms_per_step": ms, "samples_per_s": B / (ms / 1e3), "fwd_bwd_ms": ms_fb,
                         "tensor_frac_fwd_bwd": flops / (ms_fb / 1e3) / 1e12 / tf_peak, "optimizer": "torch AdamW + clip_grad_norm_(1.0), unmodified"}
    # the same step with the library's fused clip + AdamW (optim.FusedAdamW: two launches instead of torch's ~25 foreach launches)
    try:
        from osteosarcoma_diffusionmodel_b200.optim import FusedAdamW
        fopt = FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-5, max_grad_norm=1.0)

        def fstep():
            fopt.zero_grad()
            model(x0, cond, return_loss=True).backward()
            fopt.step()

        out["train_step"]["fused_optimizer_ms_per_step"] = timed(fstep, 3, 10)
        del fopt
    except Exception as e:
        out["train_step"]["fused_optimizer_error"] = repr(e)
    out["train_step"]["note"] = ("forward + backward replayed as one executable graph per batch shape; weight / bias gradients on side streams beside the dgrad chain; "
                                 "the host enqueue of torch's clip + AdamW bounds the full step")
    # multi-task step (BASELINE.json configs[3]): + pathway coherence (10 pathways x 15 genes), 2 sign rules, survival head
    try:
        from osteosarcoma_diffusionmodel_b200.multitask import BiologyConstrainedDiffusion
        rs = np.random.RandomState(0)
        members = [sorted(rs.choice(D_EXPR, 15, replace=False).tolist()) for _ in range(10)]
        mt = BiologyConstrainedDiffusion(D_MUT, D_EXPR, D_PATH, N_COND, synth.model_config(hidden_dims=HIDDEN), pathway_members=members,
                                         correlation_rules=[(0, 0, -1), (1, 1, 1)])
        mt.diffusion.load_state_dict(synth.make_params(D, N_COND, HIDDEN, seed=0), strict=False)
        mt = mt.to(dev).train()
        mt.diffusion.set_precision(model._precision)
        mopt = torch.optim.AdamW(mt.parameters(), lr=1e-4, weight_decay=1e-5)
        surv = cond[:, 0].contiguous()

        def mt_step():
            mopt.zero_grad()
            loss = mt(x0, cond, survival_time=surv)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(mt.parameters(), 1.0)
            mopt.step()

        ms_mt = timed(mt_step, 3, 10)
        out["train_step"]["multitask_ms_per_step"] = ms_mt
        out["train_step"]["multitask_parts"] = {k: float(v) for k, v in mt.last_losses.items()}
        mt.diffusion.check_status()
        del mt, mopt
    except Exception as e:      # keep the bench line if the optional workload fails
        out["train_step"]["multitask_error"] = repr(e)
    model.eval()
    # CPU port: autograd over the oracle, batch 1024
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth.make_params(D, N_COND, HIDDEN, seed=0)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    full = dict(params)
    full.update(O.schedule_buffers("cosine", T_STEPS))
    copt = torch.optim.AdamW(list(params.values()), lr=1e-4, weight_decay=1e-5)
    bx, bc = x0[:1024].cpu(), cond[:1024].cpu()
    masks = synth.dropout_masks(0, 1024, synth.block_widths(HIDDEN), 0.2)

    def cpu_step():
        copt.zero_grad()
        t = torch.randint(0, T_STEPS, (1024,))
        loss = O.forward_loss(full, bx, bc, t, torch.randn_like(bx), T_STEPS, drop_masks=masks, p=0.2, training=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(list(params.values()), 1.0)
        copt.step()

    cpu_step()
    t0 = time.perf_counter()
    for _ in range(3):
        cpu_step()
    cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    out["train_step"]["cpu_port"] = {"batch": 1024, "ms_per_step": cpu_ms, "samples_per_s": 1024 / (cpu_ms / 1e3), "cores": torch.get_num_threads()}

    # -------- standalone elementwise kernels of the path on 100k x 5142 fp32 (north_star: q_sample / reverse update at >= 70 % of HBM)
    try:
        ne = 100_000
        ex0 = torch.randn(ne, D, device=dev)
        ez = torch.randn(ne, D, device=dev)
        eeps = torch.randn(ne, D, device=dev)
        et = torch.randint(0, T_STEPS, (ne,), device=dev)
        eout = torch.empty_like(ex0)
        model._ensure_ctx(ne)
        lib_, s_ = _lib.load(), _lib.stream_handle()
        row_b = D * 4
        ecases = {
            "q_sample_injected_noise": (lambda: model.q_sample(ex0, et, ez), 3 * row_b),
            "q_sample_philox_noise": (lambda: model.q_sample(ex0, et), 3 * row_b),
            "reverse_update_injected_z": (lambda: _lib.check(lib_.osteo_ddpm_reverse_update(model._ctx, eout.data_ptr(), eeps.data_ptr(), ez.data_ptr(), ne, 500, 0, 0, s_)), 4 * row_b),
            "reverse_update_philox_z": (lambda: _lib.check(lib_.osteo_ddpm_reverse_update(model._ctx, eout.data_ptr(), eeps.data_ptr(), None, ne, 500, 7, 0, s_)), 3 * row_b),
            "store_state": (lambda: _lib.check(lib_.osteo_ddpm_store_state(model._ctx, eout.data_ptr(), ne, s_)), 2 * row_b),
        }
        out["elementwise"] = {"rows": ne, "note": "algorithmic bytes (each tensor read or written once) / CUDA-event time, through the public calls"}
        for name, (fn, bpr) in ecases.items():
            ms = timed(fn, 2, 5)
            gbs = bpr * ne / (ms / 1e3) / 1e9
            out["elementwise"][name] = {"ms": ms, "gb_per_s": gbs, "hbm_frac": gbs / hbm_peak}
        del ex0, ez, eeps, eout
    except Exception as e:
        out["elementwise"] = {"error": repr(e)}

    # -------- training ingress: one 8192-row batch gathered from a device-resident cohort with the mixup fused in
    try:
        from osteosarcoma_diffusionmodel_b200.ingress import GpuResidentDataset, MixupAugmentation
        nd = 32768
        dsx, dsc = synth.make_cohort(nd, D_MUT, D_EXPR, D_PATH, N_COND, seed=5)
        ds = GpuResidentDataset(dsx, dsc, dsc[:, 0].contiguous(), device=dev)
        gi = torch.Generator(device=dev).manual_seed(0)
        index = torch.randperm(nd, device=dev, generator=gi)[:8192].contiguous()
        perm = torch.randperm(8192, device=dev, generator=gi)
        mixer = MixupAugmentation(0.2)
        ms = timed(lambda: mixer.gather(ds, index, lam=0.3, perm=perm), 2, 10)
        moved = 8192 * (D + N_COND + 1) * 4 * 3
        host_batch = {"data": dsx[:8192].clone(), "conditions": dsc[:8192].clone()}
        ms_h2d = timed(lambda: (host_batch["data"].to(dev), host_batch["conditions"].to(dev)), 1, 5)
        out["ingress"] = {"batch": 8192, "dataset_rows": nd, "fused_gather_mixup_ms": ms, "gb_per_s": moved / (ms / 1e3) / 1e9, "hbm_frac": moved / (ms / 1e3) / 1e9 / hbm_peak,
                          "reference_style_pageable_h2d_ms": ms_h2d,
                          "note": "read two dataset rows + write one per batch row (data, conditions, survival); the reference copies each batch from pageable host memory (utils/train.py:214-216) and mixes with three elementwise passes"}
        del ds, dsx
    except Exception as e:
        out["ingress"] = {"error": repr(e)}

    # -------- RBF-MMD (utils/validation.py:273-298): N = M = 16384 rows of 5142 features
    n = 16384
    g = torch.Generator(device=dev).manual_seed(1)
    X = torch.randn(n, D, device=dev, generator=g) + 4.0
    Y = torch.randn(n, D, device=dev, generator=g) * 1.1 + 4.1
    for prec in ("bf16", "fp32x3"):
        val = BiologicalValidator({"evaluation": {}}, precision=prec)
        ms = timed(lambda: val.compute_mmd(X, Y), 1, 3)
        pairs = 3.0 * n * n
        passes = 3 if prec == "fp32x3" else 1
        # half-Gram for XX and YY: 2 * n(n+1)/2 + n*n tile-pairs actually multiplied
        mults = (n * (n + 128) + n * n) * 2.0 * D * passes
        out[f"mmd_{prec}"] = {"rows": n, "ms": ms, "kernel_pairs_per_s": pairs / (ms / 1e3), "tensor_tflops": mults / (ms / 1e3) / 1e12,
                              "tensor_frac": mults / (ms / 1e3) / 1e12 / tf_peak, "value": val.compute_mmd(X, Y), "includes": "centring, bf16 packing, row norms, 3 Gram reductions"}
    Xc, Yc = X[:384].cpu().numpy(), Y[:384].cpu().numpy()
    t0 = time.perf_counter()
    V.compute_mmd(Xc, Yc)
    dt = time.perf_counter() - t0
    out["mmd_cpu_port"] = {"rows": 384, "s": dt, "kernel_pairs_per_s": 3 * 384 * 384 / dt, "note": "numpy restatement of scipy cdist + exp, literal differences"}

    # -------- pathway coherence moments (utils/validation.py:125-175): 1M rows, 10 pathways x 15 genes
    rows = 1_000_000
    cohort = torch.randn(rows, 371, device=dev, generator=g)
    members = [list(range(15 * p, 15 * p + 15)) for p in range(10)]
    val = BiologicalValidator({"evaluation": {}})
    real_c, syn_c = cohort[: rows // 2], cohort[rows // 2:]
    ms = timed(lambda: val.pathway_coherence_from_tensors(real_c, syn_c, members), 2, 10)
    # the device work of that call alone (two moment kernels + the Pearson finish, enqueued back to back, no host synchronisation inside)
    from osteosarcoma_diffusionmodel_b200.validation import _CM_STRIDE, _coherence_finish
    mom = torch.empty((2 * len(members), _CM_STRIDE), dtype=torch.float64, device=dev)
    ci_pair = torch.cat([val._index_tensor(dev, members), val._index_tensor(dev, members)])

    def coherence_device_only():
        val._coherence_moments(real_c, members, out=mom[: len(members)], reduce=False)
        val._coherence_moments(syn_c, members, out=mom[len(members):], reduce=False)
        _coherence_finish(mom, ci_pair)

    ms_dev = timed(coherence_device_only, 2, 10)
    out["coherence"] = {"rows": rows, "genes": 371, "pathways": 10, "ms": ms, "gathered_gb_per_s": rows * 150 * 4 / (ms / 1e3) / 1e9,
                        "streamed_gb_per_s": rows * 371 * 4 / (ms / 1e3) / 1e9, "hbm_frac": rows * 371 * 4 / (ms / 1e3) / 1e9 / hbm_peak,
                        "device_ms": ms_dev, "device_hbm_frac": rows * 371 * 4 / (ms_dev / 1e3) / 1e9 / hbm_peak,
                        "note": "both cohorts: register-tiled moment kernel (osteo_corr_moments_tiled, one bulk copy per 32-row chunk, one pass over whole rows for all pathways) + Pearson finish on the device + one 20-double D2H; ms = the whole public call (host work and the D2H synchronisation included), device_ms = the two moment kernels + finish alone"}
    del real_c, syn_c
    del cohort, X, Y

    # -------- fp32x3 (split-bf16, fp32-tolerance) sampling throughput beside the bf16 headline: same workload, one full loop
    try:
        nrows = 100_000
        cond_dev = synth.scenario_conditions(nrows, N_COND).to(dev)
        model.set_precision("fp32x3")
        model.sample(cond_dev, nrows, seed=1, t_stop=T_STEPS - 20)
        ms = timed(lambda: model.sample(cond_dev, nrows, seed=2), 0, 1)
        out["sampling_fp32x3"] = {"rows": nrows, "ms": ms, "patients_per_s": nrows / (ms / 1e3), "tolerance_vs_reference": "rel 1e-4 (tests/helpers.py TOL_FP32X3)"}
    except Exception as e:
        out["sampling_fp32x3"] = {"error": repr(e)}
    finally:
        model.set_precision("bf16")

    # -------- the reference itself on THIS GPU under PyTorch eager (SURVEY.md §2.1 bar (i)): fp32 cuBLAS + ~45 launches per step
    try:
        ref = reference_model(dev)
        if ref is None:
            out["gpu_eager_reference"] = {"unavailable": "no staged reference (oracle/_ref/reference)"}
        else:
            out["gpu_eager_reference"] = {}
            for nrows in (1024, 32768):
                cond_r = synth.scenario_conditions(nrows, N_COND).to(dev)
                state = {"x": torch.randn(nrows, D, device=dev), "t": T_STEPS - 1}

                def ref_step():
                    with torch.no_grad():
                        state["x"] = ref.p_sample(state["x"], state["t"], cond_r)
                    state["t"] = state["t"] - 1 if state["t"] > 1 else T_STEPS - 1

                ms = timed(ref_step, 5, 40)
                out["gpu_eager_reference"][f"rows_{nrows}"] = {"ms_per_reverse_step": ms, "patients_per_s": nrows / (ms * T_STEPS / 1e3)}
            out["gpu_eager_reference"]["note"] = ("the staged reference's own p_sample (models/diffusion.py:382-425) on this B200, PyTorch eager fp32 (TF32 off, torch default), "
                                                  "40 reverse steps scaled to 1000")
            del ref
    except Exception as e:
        out["gpu_eager_reference"] = {"error": repr(e)}
    return out


def run_secondary_dp(model, dev, dist, rank, world, tf_peak):
    """N > 1: the data-parallel training step of BASELINE.json configs[3] (8192 rows per GPU, one all-reduce of the 17 MB flat gradient
    buffer, clip after the reduce, unmodified torch AdamW) and the row-sharded RBF-MMD of configs[4] (Gram rows split over the ranks,
    three fp64 sums all-reduced). Device-timed between barriers, max over ranks."""
    import torch
    from osteosarcoma_diffusionmodel_b200 import distributed as Dm
    from osteosarcoma_diffusionmodel_b200 import synthetic as synth
    from osteosarcoma_diffusionmodel_b200.validation import BiologicalValidator

    def timed(fn, warm, reps):
        for _ in range(warm):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    out = {}
    B = 8192
    x0, cond = synth.make_cohort(B, D_MUT, D_EXPR, D_PATH, N_COND, seed=3 + rank)
    x0, cond = x0.to(dev), cond.to(dev)
    model.train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)
    ms = timed(lambda: Dm.dp_train_step(model, opt, x0, cond), 4, 10)
    out["dp_train_step"] = {"batch_per_gpu": B, "global_batch": B * world, "ms_per_step": ms, "samples_per_s": B * world / (ms / 1e3),
                            "allreduce": "one NCCL all-reduce of the flat fp32 gradient buffer per step, after the graph-replayed backward; *_overlap = the "
                                         "backward pass cut in two graph launches with the all-reduce of the first part's gradients (output_proj + decoder, "
                                         "about half of the 17 MB) running beside the second part (model._dp_overlap / OSTEO_DP_OVERLAP=1)"}
    try:      # the same step with the library's fused clip + AdamW (two launches; the torch optimiser's host enqueue bounds the step above)
        from osteosarcoma_diffusionmodel_b200.optim import FusedAdamW
        fopt = FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-5, max_grad_norm=1.0)
        ms_f = timed(lambda: Dm.dp_train_step(model, fopt, x0, cond), 4, 10)
        out["dp_train_step"]["fused_optimizer_ms_per_step"] = ms_f
        # configs[3]: "report step time with and without allreduce overlap"
        keep = model._dp_overlap
        model._dp_overlap = True
        out["dp_train_step"]["fused_optimizer_ms_per_step_overlap"] = timed(lambda: Dm.dp_train_step(model, fopt, x0, cond), 4, 10)
        model._dp_overlap = keep
        del fopt
    except Exception as e:
        out["dp_train_step"]["fused_optimizer_error"] = repr(e)
    model.eval()
    n = 32768
    g = torch.Generator(device=dev).manual_seed(1)
    X = torch.randn(n, D, device=dev, generator=g) + 4.0
    Y = torch.randn(n, D, device=dev, generator=g) * 1.1 + 4.1
    val = BiologicalValidator({"evaluation": {}}, precision="bf16")
    ms = timed(lambda: val.compute_mmd(X, Y), 1, 3)
    out["mmd_bf16_row_sharded"] = {"rows": n, "ms": ms, "kernel_pairs_per_s": 3.0 * n * n / (ms / 1e3), "value": val.compute_mmd(X, Y),
                                   "note": "every rank holds X and Y and reduces the Gram row blocks b % world == rank (block-cyclic), Kxx and Kyy as symmetric half-Grams"}
    # ---- correctness of every sharded path, outside any timed region, reported in the JSON line (secondary.checks)
    out["checks"] = multi_gpu_checks(model, dev, dist, rank, world, opt, X, Y)
    model.check_status()
    return out if rank == 0 else None


def multi_gpu_checks(model, dev, dist, rank, world, opt, X, Y):
    """Driver-visible multi-GPU correctness (every rank takes part; all comparisons are all-reduced so every rank agrees on the verdict):
      sampling_shards_equal_single_gpu   rows sampled by rank r inside its shard == the same GLOBAL rows sampled by rank 0 alone (bit-equal)
      dp_replicas_bit_identical          parameters (and AdamW moments) identical on all ranks after data-parallel optimiser steps
      mmd_sharded_equals_unsharded       row-sharded RBF-MMD == the single-GPU reduction of the same X, Y (relative 1e-6)
      coherence_sharded_equals_unsharded row-sharded pathway-coherence scores == the single-GPU ones (absolute 1e-7)"""
    import torch
    from osteosarcoma_diffusionmodel_b200 import distributed as Dm
    from osteosarcoma_diffusionmodel_b200 import synthetic as synth
    from osteosarcoma_diffusionmodel_b200 import validation as Vm
    from osteosarcoma_diffusionmodel_b200.validation import BiologicalValidator

    checks = {}

    def agree(ok: bool) -> bool:
        t = torch.tensor([1 if ok else 0], device=dev, dtype=torch.int32)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    # -- sampling: full 1000-step loop, a global cohort of world x 384 patients (three row tiles per rank, ragged against 128)
    model.eval()
    per, seed = 384 - 7, 4321
    n_glob = per * world
    cond_glob = synth.scenario_conditions(n_glob, N_COND).to(dev)
    local = Dm.sample_sharded(model, cond_glob, n_glob, seed=seed)                     # this rank's contiguous slice
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    ok = True
    if rank == 0:
        whole = model.sample(cond_glob, n_glob, seed=seed, row_base=0)                  # all global rows on ONE GPU
        ok = bool(torch.equal(whole, torch.cat(gathered))) and bool(torch.isfinite(whole).all())
    checks["sampling_shards_equal_single_gpu"] = agree(ok)
    checks["sampling_rows_checked"] = n_glob

    # -- data-parallel training: after the steps timed above plus two more, every replica holds the same bits
    model.train()
    x0, cond = synth.make_cohort(2048, D_MUT, D_EXPR, D_PATH, N_COND, seed=100 + rank)
    for _ in range(2):
        Dm.dp_train_step(model, opt, x0.to(dev), cond.to(dev))
    model.eval()
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    state = [v for st in opt.state.values() for k, v in sorted(st.items()) if torch.is_tensor(v) and v.numel() > 1]
    flat = torch.cat([flat] + [v.detach().reshape(-1).float() for v in state])
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    checks["dp_replicas_bit_identical"] = agree(bool(torch.equal(ref.view(torch.int32), flat.view(torch.int32))) and bool(torch.isfinite(flat).all()))

    # -- RBF-MMD: sharded Gram rows + all-reduce vs the whole reduction on one GPU
    for prec in ("bf16", "fp32x3"):
        val = BiologicalValidator({"evaluation": {}}, precision=prec)
        n = 4096 + 100
        sharded = val.compute_mmd(X[:n], Y[:n - 333])
        center = ((X[:n].sum(0, dtype=torch.float64) + Y[:n - 333].sum(0, dtype=torch.float64)) / (2 * n - 333)).float().contiguous()
        sums = Vm._gram_partial_sums(X[:n].contiguous(), Y[:n - 333].contiguous(), 1.0 / D, center, (0, n), (0, n - 333), Vm._PRECISIONS[prec]).cpu()
        sxx, syy, sxy = (float(v) for v in sums)
        whole = max(sxx / n ** 2 + syy / (n - 333) ** 2 - 2 * sxy / (n * (n - 333.0)), 0.0) ** 0.5
        checks[f"mmd_sharded_equals_unsharded_{prec}"] = agree(abs(sharded - whole) <= 1e-6 * max(whole, 1e-12))
    checks["mmd_sharded_equals_unsharded"] = checks["mmd_sharded_equals_unsharded_bf16"] and checks["mmd_sharded_equals_unsharded_fp32x3"]

    # -- pathway coherence: sharded cohort rows + all-reduced moment blocks vs one GPU
    g = torch.Generator(device=dev).manual_seed(7)
    cohort = torch.randn(50_001, 371, device=dev, generator=g)
    cohort[:, 1] += 0.5 * cohort[:, 0]
    members = [list(range(15 * p, 15 * p + 15)) for p in range(10)]
    val = BiologicalValidator({"evaluation": {}})
    sharded = val._coherence_scores(cohort, members)
    packs_key = (str(cohort.device), tuple(tuple(c) for c in members))
    ci_t, gather_idx = val._index_cache[packs_key][0]
    mom = Vm._moments_tiled(cohort, ci_t, (0, cohort.shape[0]), 15).cpu().numpy()      # the same kernel over ALL rows on this GPU alone
    whole = val._scores_from_moments(mom, members)
    # fp32 partial sums over <= 4 chunks of 32 rows, fp64 across chunks: a different row partition changes the fp32 roundings -- scores
    # agree to ~1e-8; the validators' stated tolerance is 1e-6 absolute (tests/test_validators_gpu.py)
    diff = max(abs(a - b) for a, b in zip(sharded, whole))
    checks["coherence_sharded_equals_unsharded"] = agree(diff <= 1e-7)
    checks["coherence_max_abs_diff"] = diff
    return checks


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--rows", type=int, default=100_000, help="patients per GPU (BASELINE.json configs[1]: 100k)")
    ap.add_argument("--precision", choices=["bf16", "fp32x3"], default="bf16")
    ap.add_argument("--chunk-rows", type=int, default=None)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the train-step / MMD / coherence timings")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3          # timing rule: at least 3 warm-up steps
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
