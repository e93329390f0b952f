"""Diagnostic: cProfile of the host side of 200 fwd+bwd calls at batch 8192 (the step is host-bound once the GPU work is one graph)."""
import cProfile, pstats, sys
sys.path.insert(0, ".")
import torch
from osteosarcoma_diffusionmodel_b200 import synthetic as synth
from osteosarcoma_diffusionmodel_b200.diffusion import BiologyAwareDiffusionModel

B = 8192
model = BiologyAwareDiffusionModel(62, 5054, 26, 3, synth.model_config())
model.load_state_dict(synth.make_params(5142, 3, (256, 512, 256), seed=0), strict=False)
model = model.to("cuda").train()
x0, cond = synth.make_cohort(B, 62, 5054, 26, 3, seed=3)
x0, cond = x0.cuda(), cond.cuda()
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)
which = sys.argv[1] if len(sys.argv) > 1 else "fwd_bwd"


def fwd_bwd():
    model.zero_grad()
    model(x0, cond, return_loss=True).backward()


def step():
    opt.zero_grad()
    loss = model(x0, cond, return_loss=True)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()


fn = fwd_bwd if which == "fwd_bwd" else step
for _ in range(5):
    fn()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    fn()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
