"""Short, ncu-friendly run: 3 reverse steps of the bench workload (100k patients, config.yaml dims, bf16) launched
eagerly (no graph) so every kernel is a separate launch. Used for profiles/*.csv; never a bench number."""
import sys
sys.path.insert(0, ".")
import torch
from osteosarcoma_diffusionmodel_b200 import synthetic as synth
from osteosarcoma_diffusionmodel_b200.diffusion import BiologyAwareDiffusionModel

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
model = BiologyAwareDiffusionModel(62, 5054, 26, 3, synth.model_config())
model.load_state_dict(synth.make_params(5142, 3, (256, 512, 256), seed=0), strict=False)
model = model.to("cuda").eval()
model._use_graph = False
cond = synth.scenario_conditions(rows, 3).cuda()
out = model.sample(cond, rows, seed=1, t_stop=1000 - steps)
torch.cuda.synchronize()
model.check_status()
print("ok", out.shape, float(out.abs().mean()))
