"""Diagnostic: kernel timeline (start, duration, stream) of one fwd+bwd call at batch 8192 from torch.profiler (CUPTI)."""
import sys
sys.path.insert(0, ".")
import torch
from torch.profiler import profile, ProfilerActivity
from osteosarcoma_diffusionmodel_b200 import synthetic as synth
from osteosarcoma_diffusionmodel_b200.diffusion import BiologyAwareDiffusionModel

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
model = BiologyAwareDiffusionModel(62, 5054, 26, 3, synth.model_config())
model.load_state_dict(synth.make_params(5142, 3, (256, 512, 256), seed=0), strict=False)
model = model.to("cuda").train()
x0, cond = synth.make_cohort(B, 62, 5054, 26, 3, seed=3)
x0, cond = x0.cuda(), cond.cuda()


def fwd_bwd():
    model.zero_grad()
    model(x0, cond, return_loss=True).backward()


for _ in range(5):
    fwd_bwd()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        fwd_bwd()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# last call only
starts = [e for e in evs if "train_prepare" in e.name]
t0 = starts[-1].time_range.start
print(f"{'start_us':>9} {'dur_us':>8} {'stream':>6}  kernel")
for e in evs:
    if e.time_range.start >= t0:
        print(f"{(e.time_range.start - t0):9.1f} {e.time_range.elapsed_us():8.1f} {getattr(e, 'stream', -1) if hasattr(e, 'stream') else -1:>6}  {e.name[:110]}")
