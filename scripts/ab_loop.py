"""Same-box A/B of the full sampling loop: run with OSTEO_DDPM_LIB pointing at each library in turn (separate processes, alternating),
because box-to-box differences (power cap, +-3 %) are as large as the effects being measured. Prints ms per reverse step of the timed
1000-step loop at 100k patients after one warm-up loop."""
import sys, time
sys.path.insert(0, ".")
import torch
from osteosarcoma_diffusionmodel_b200 import synthetic as synth
from osteosarcoma_diffusionmodel_b200.diffusion import BiologyAwareDiffusionModel

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
loops = int(sys.argv[2]) if len(sys.argv) > 2 else 2
model = BiologyAwareDiffusionModel(62, 5054, 26, 3, synth.model_config())
model.load_state_dict(synth.make_params(5142, 3, (256, 512, 256), seed=0), strict=False)
model = model.to("cuda").eval()
cond = synth.scenario_conditions(rows, 3).cuda()
model.sample(cond, rows, seed=1)          # warm-up loop (graph capture, clocks settle under the power cap)
torch.cuda.synchronize()
for i in range(loops):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = model.sample(cond, rows, seed=2 + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"loop {i}: {ms / 1000:.4f} ms/step  {rows / ms * 1000 / 1000:.1f} patients/s (x1000 steps)", flush=True)
model.check_status()
