"""GPU: the data-parallel step with the backward pass cut in two and the gradient all-reduce of the first part started in between
(osteo_ddpm_train_backward_part) gives the gradients of the one-launch step. One rank over NCCL (a 1-GPU box can run it); the 2-rank
behaviour is covered by tests/test_multigpu_gpu.py, which switches the overlap on."""
import os
import socket

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


def _worker(port, q):
    import torch.distributed as dist
    from oracle import synth
    from osteosarcoma_diffusionmodel_b200 import distributed as D
    from tests.helpers import build_model, load_case, rel

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK="0", WORLD_SIZE="1")
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    res = {}
    for name in ("smoke", "linear3"):
        case = load_case(name)
        d = case["dims"]
        x0, c = synth.make_cohort(256, d["mutation_dim"], d["expression_dim"], d["pathway_dim"], d["condition_dim"], seed=6)
        x0, c = x0.cuda(), c.cuda()
        runs = {}
        for overlap in (False, True):
            model = build_model(case, "bf16")
            model.train()
            model._dp_overlap = overlap
            model._flat_allreduce = True
            # world size 1 would skip the all-reduce path altogether: force the two code paths by hand
            out = []
            for step in range(3):          # step 0 eager, 1 captures the graphs, 2 replays them
                model.manual_seed(100 + step)          # in-kernel noise / dropout streams
                t = torch.randint(0, case["T"], (x0.shape[0],), generator=torch.Generator().manual_seed(step))
                loss, grads = model._run_train_step(x0, c, {"t": t}, want_grads=True, dp_overlap=overlap)
                out.append((float(loss), [g.detach().clone() for g in grads]))
            model.check_status()
            runs[overlap] = out
            cut, off = model._dp_cut(model._param_list())
            res[f"{name}_cut"] = (cut, off, sum(p.numel() for p in model._param_list()))
        worst = 0.0
        for (la, ga), (lb, gb) in zip(runs[False], runs[True]):
            assert abs(la - lb) <= 2e-5 * abs(la), (la, lb)
            for a, b in zip(ga, gb):
                worst = max(worst, min(rel(b, a), float((a - b).abs().max())))
        res[name] = worst
    # the public data-parallel step with the overlap on
    case = load_case("linear3")
    model = build_model(case, "bf16")
    model.train()
    model._dp_overlap = True
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3)
    x0, c = synth.make_cohort(128, 20, 90, 10, 2, seed=9)
    losses = [float(D.dp_train_step(model, opt, x0.cuda(), c.cuda())) for _ in range(3)]
    res["dp_losses_finite"] = all(l == l and l < 1e6 for l in losses)
    q.put(res)
    dist.destroy_process_group()


def test_split_backward_matches_the_one_launch_step():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_worker, args=(_free_port(), q))
    p.start()
    p.join(600)
    assert p.exitcode == 0
    res = q.get(timeout=10)
    # split-batch weight gradients accumulate atomically: equal to rounding noise, not bit for bit (tests/test_training_gpu.py)
    assert res["smoke"] < 1e-4 and res["linear3"] < 1e-4, res
    for name in ("smoke", "linear3"):
        cut, off, total = res[f"{name}_cut"]
        assert 0 < off < total and 0.25 < off / total < 0.75, res
    assert res["dp_losses_finite"]
