"""Diagnostic: fwd+bwd time of the training step (batch argv[1], default 8192) and the host-side enqueue time of one call."""
import sys, time
sys.path.insert(0, ".")
import torch
from osteosarcoma_diffusionmodel_b200 import synthetic as synth
from osteosarcoma_diffusionmodel_b200.diffusion import BiologyAwareDiffusionModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
model = BiologyAwareDiffusionModel(62, 5054, 26, 3, synth.model_config())
model.load_state_dict(synth.make_params(5142, 3, (256, 512, 256), seed=0), strict=False)
model = model.to("cuda").train()
x0, cond = synth.make_cohort(B, 62, 5054, 26, 3, seed=3)
x0, cond = x0.cuda(), cond.cuda()
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)


def fwd_bwd():
    model.zero_grad()
    model(x0, cond, return_loss=True).backward()


def step():
    opt.zero_grad()
    loss = model(x0, cond, return_loss=True)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()


from osteosarcoma_diffusionmodel_b200.optim import FusedAdamW
fopt = FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-5, max_grad_norm=1.0)


def fused_step():
    fopt.zero_grad()
    model(x0, cond, return_loss=True).backward()
    fopt.step()


for name, fn in (("fwd_bwd", fwd_bwd), ("full step", step), ("full step, FusedAdamW", fused_step)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(20):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / 20 * 1e3
    torch.cuda.synchronize()
    print(f"batch {B} {name}: {e0.elapsed_time(e1) / 20:.3f} ms on the device, {host:.3f} ms host enqueue", flush=True)
model.check_status()

# multi-task step: 10 pathways x 15 genes, 2 sign rules, survival head
import numpy as np
from osteosarcoma_diffusionmodel_b200.multitask import BiologyConstrainedDiffusion
rs = np.random.RandomState(0)
mt = BiologyConstrainedDiffusion(62, 5054, 26, 3, synth.model_config(), pathway_members=[sorted(rs.choice(5054, 15, replace=False).tolist()) for _ in range(10)],
                                 correlation_rules=[(0, 0, -1), (1, 1, 1)])
mt.diffusion.load_state_dict(synth.make_params(5142, 3, (256, 512, 256), seed=0), strict=False)
mt = mt.to("cuda").train()
mopt = FusedAdamW(mt.parameters(), lr=1e-4, weight_decay=1e-5, max_grad_norm=1.0)
surv = cond[:, 0].contiguous()


def mt_step():
    mopt.zero_grad()
    mt(x0, cond, survival_time=surv).backward()
    mopt.step()


for _ in range(3):
    mt_step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(20):
    mt_step()
e1.record()
host = (time.perf_counter() - t0) / 20 * 1e3
torch.cuda.synchronize()
print(f"batch {B} multi-task step, FusedAdamW: {e0.elapsed_time(e1) / 20:.3f} ms on the device, {host:.3f} ms host", flush=True)
if len(sys.argv) > 2:
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(50):
        mt_step()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
