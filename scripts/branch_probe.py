"""Diagnostic: 1000-step sampling time against the number of parallel row branches of the sampling graph (and the chunk size),
plus a bit-exactness check of the samples across settings."""
import sys, time
sys.path.insert(0, ".")
import torch
from osteosarcoma_diffusionmodel_b200 import synthetic as synth
from osteosarcoma_diffusionmodel_b200.diffusion import BiologyAwareDiffusionModel

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
settings = [tuple(int(v) for v in a.split(":")) for a in sys.argv[2:]] or [(1, 131072), (2, 131072), (3, 131072), (4, 131072)]
model = BiologyAwareDiffusionModel(62, 5054, 26, 3, synth.model_config())
model.load_state_dict(synth.make_params(5142, 3, (256, 512, 256), seed=0), strict=False)
model = model.to("cuda").eval()
cond = synth.scenario_conditions(rows, 3).cuda()
ref = None
for nb, chunk in settings:
    model.set_branches(nb)
    model.set_chunk_rows(chunk)
    model.sample(cond, rows, seed=1, t_stop=990)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    best = 1e9
    for rep in range(2):
        torch.cuda.synchronize()
        ev[0].record()
        out = model.sample(cond, rows, seed=7)
        ev[1].record()
        torch.cuda.synchronize()
        best = min(best, ev[0].elapsed_time(ev[1]))
    chk = out[::997].clone()
    same = True if ref is None else bool(torch.equal(chk, ref))
    if ref is None:
        ref = chk
    print(f"branches={nb} chunk={chunk}: {best:.1f} ms per 1000 steps, {rows / best * 1e3:.0f} patients/s, bit-identical to first setting: {same}", flush=True)
    del out
