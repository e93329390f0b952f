"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): the multi-GPU paths of SURVEY.md §8(e) end to end."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, q):
    import torch.distributed as dist
    from oracle import synth
    from osteosarcoma_diffusionmodel_b200 import distributed as D
    from osteosarcoma_diffusionmodel_b200.validation import BiologicalValidator
    from tests.helpers import build_model, load_case, rel

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=ws, device_id=torch.device("cuda", rank))
    dev = f"cuda:{rank}"
    case = load_case("linear3")
    res = {}
    # ---- batch-sharded sampling: no collective on the data path, union identical to the single-GPU cohort
    model = build_model(case, "bf16", device=dev)
    n = 300
    _, cond = synth.make_cohort(n, 20, 90, 10, 2, seed=4)
    cond = cond.to(dev)
    full = D.sample_sharded(model, cond, n, seed=5, gather=True)
    single = model.sample(cond, n, seed=5)          # every rank also computes the whole cohort locally
    res["sample_equal"] = bool(torch.equal(full, single))
    # ---- data-parallel training: replicas stay identical and match one big-batch step
    model = build_model(case, "fp32x3", device=dev)
    ref = build_model(case, "fp32x3", device=dev)
    model.train(); ref.train()
    B = 128
    x0, c = synth.make_cohort(B, 20, 90, 10, 2, seed=9)
    rs = np.random.RandomState(1)
    t = torch.from_numpy(rs.randint(0, case["T"], size=B).astype(np.int64))
    noise = synth.noise_stream(3)(1, (B, case["D"]))
    masks = synth.dropout_masks(3, B, synth.block_widths(case["hidden"]), 0.2)
    lo, hi = D.shard_rows(B, rank, ws)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    model._inject = {"t": t[lo:hi], "noise": noise[lo:hi], "masks": [m[lo:hi] for m in masks]}
    model._dp_overlap = True          # the two-piece all-reduce with the backward pass cut in between (opt-in): same result
    D.dp_train_step(model, opt, x0[lo:hi].to(dev), c[lo:hi].to(dev))
    ropt = torch.optim.AdamW(ref.parameters(), lr=1e-3)
    ref._inject = {"t": t, "noise": noise, "masks": masks}
    ropt.zero_grad()
    ref(x0.to(dev), c.to(dev)).backward()
    torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
    ropt.step()
    res["dp_vs_big_batch"] = max(rel(p, r) for p, r in zip(model.parameters(), ref.parameters()))
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    other = [torch.zeros_like(flat) for _ in range(ws)]
    dist.all_gather(other, flat)
    res["replicas_identical"] = all(bool(torch.equal(other[0], o)) for o in other)
    # ---- row-sharded MMD + coherence
    rs = np.random.RandomState(2)
    X = (rs.standard_normal((700, 96)) + 3).astype(np.float32)
    Y = (rs.standard_normal((520, 96)) * 1.1 + 3.2).astype(np.float32)
    val = BiologicalValidator({"evaluation": {}}, device=dev)
    res["mmd"] = val.compute_mmd(X, Y)
    res["coh"] = val.pathway_coherence_from_tensors(torch.from_numpy(X), torch.from_numpy(Y[:, :96]), [[0, 1, 2, 3], [10, 11, 12], [20, 30, 40, 50, 60]])
    # ---- data-parallel correlation losses (multi-task step): every rank holds a row shard, the moments are all-reduced, so the loss is
    # the GLOBAL batch's and each rank's gradient is world_size x its rows of the global gradient (DP averaging divides it back)
    from osteosarcoma_diffusionmodel_b200.multitask import correlation_losses
    rs = np.random.RandomState(7)
    G = torch.from_numpy((rs.standard_normal((600, 3)) @ rs.standard_normal((3, 24)) + rs.standard_normal((600, 24))).astype(np.float32))
    sets, modes = [[0, 1, 2, 3, 4], [5, 6, 7], [8, 9], [10, 11]], [0, 0, 1, -1]
    lo, hi = D.shard_rows(600, rank, ws)
    xs = G[lo:hi].to(dev).requires_grad_(True)
    ls = correlation_losses(xs, sets, modes)
    ls.sum().backward()
    xf = G.double().requires_grad_(True)
    from oracle import bio_losses_oracle as Bo
    lf = Bo.correlation_losses(xf, sets, modes)
    lf.sum().backward()
    res["corr_loss_err"] = float((ls.detach().cpu().double() - lf.detach()).abs().max())
    res["corr_grad_err"] = rel(xs.grad.cpu(), ws * xf.grad[lo:hi])
    errs = torch.tensor([res["corr_loss_err"], res["corr_grad_err"]], device=dev, dtype=torch.float64)
    dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    res["corr_loss_err"], res["corr_grad_err"] = errs.tolist()
    if rank == 0:
        q.put(res)
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_paths():
    from oracle import validators_oracle as V

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(600) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    res = q.get(timeout=10)
    assert res["sample_equal"]
    assert res["replicas_identical"]
    assert res["dp_vs_big_batch"] < 2e-4
    rs = np.random.RandomState(2)
    X = (rs.standard_normal((700, 96)) + 3).astype(np.float32)
    Y = (rs.standard_normal((520, 96)) * 1.1 + 3.2).astype(np.float32)
    assert abs(res["mmd"] - V.compute_mmd(X, Y)) < 1e-4 * V.compute_mmd(X, Y)
    ref = V.pathway_coherence(X, Y, [[0, 1, 2, 3], [10, 11, 12], [20, 30, 40, 50, 60]])
    for k in ref:
        assert abs(res["coh"][k] - ref[k]) < 1e-6
    assert res["corr_loss_err"] < 5e-6 and res["corr_grad_err"] < 5e-5


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_model_on_a_non_current_device():
    """a model on cuda:1 while the process's current device is cuda:0 -- every C-ABI call runs under a device guard and on the
    MODEL device's current stream; the current device is unchanged afterwards (construction, sampling, training, garbage collection)."""
    import gc
    from oracle import synth
    from tests.helpers import build_model, load_case

    torch.cuda.set_device(0)
    case = load_case("linear3")
    ref_model = build_model(case, "bf16", device="cuda:0")
    model = build_model(case, "bf16", device="cuda:1")
    n = 200
    _, cond = synth.make_cohort(n, 20, 90, 10, 2, seed=4)
    a = ref_model.sample(cond.to("cuda:0"), n, seed=5, t_stop=990)
    b = model.sample(cond.to("cuda:1"), n, seed=5, t_stop=990)
    assert torch.cuda.current_device() == 0
    assert b.device == torch.device("cuda", 1) and torch.equal(a.cpu(), b.cpu())
    model.train()
    x0, c = synth.make_cohort(64, 20, 90, 10, 2, seed=2)
    loss = model(x0.to("cuda:1"), c.to("cuda:1"))
    loss.backward()
    assert torch.isfinite(loss).item() and torch.cuda.current_device() == 0
    model.check_status()
    # the optimiser step and the validators too: their kernels opt in to large shared memory PER DEVICE (a process-wide "configured" flag
    # made the first launch on cuda:1 fail with "invalid argument" after cuda:0 had been used)
    from osteosarcoma_diffusionmodel_b200.optim import FusedAdamW
    from osteosarcoma_diffusionmodel_b200.validation import BiologicalValidator
    opt = FusedAdamW(model.parameters(), lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
    opt.step()
    assert torch.cuda.current_device() == 0
    g = torch.Generator().manual_seed(3)
    X, Y = torch.randn(300, 120, generator=g), torch.randn(260, 120, generator=g) + 0.1
    members = [list(range(12 * i, 12 * i + 12)) for i in range(4)]
    res = {}
    for dev in ("cuda:0", "cuda:1"):
        val = BiologicalValidator({"evaluation": {}}, device=dev)
        res[dev] = (val.compute_mmd(X.to(dev), Y.to(dev)), val.pathway_coherence_from_tensors(X.to(dev), Y[:, :120].to(dev), members))
        assert torch.cuda.current_device() == 0
    assert res["cuda:0"][0] == res["cuda:1"][0]
    for k, v in res["cuda:0"][1].items():
        assert abs(v - res["cuda:1"][1][k]) < 1e-9
    del model
    gc.collect()
    assert torch.cuda.current_device() == 0
