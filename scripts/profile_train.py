"""ncu-friendly: 3 training steps (fwd + bwd + clip + AdamW) at batch 8192, config.yaml dims, bf16. Not a bench number."""
import sys
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
for _ in range(3):
    opt.zero_grad(set_to_none=False)
    loss = model(x0, cond)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()
torch.cuda.synchronize()
model.check_status()
print("ok", float(loss))
