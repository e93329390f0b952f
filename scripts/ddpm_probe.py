"""Diagnostic: per-kernel times of one reverse step at t=500 (Philox noise) vs t=0 (sigma = 0: no RNG at all)."""
import sys, ctypes as C
sys.path.insert(0, ".")
import torch
from osteosarcoma_diffusionmodel_b200 import synthetic as synth
from osteosarcoma_diffusionmodel_b200.diffusion import BiologyAwareDiffusionModel
from osteosarcoma_diffusionmodel_b200 import _lib

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
model = BiologyAwareDiffusionModel(62, 5054, 26, 3, synth.model_config())
model.load_state_dict(synth.make_params(5142, 3, (256, 512, 256), seed=0), strict=False)
model = model.to("cuda").eval()
cond = synth.scenario_conditions(rows, 3).cuda()
model.sample(cond, rows, seed=1, t_stop=998)
lib = _lib.load()
buf = (C.c_float * 4096)()
import os
for t, dbg in [(500, int(d)) for d in os.environ.get('PROBE_DBG', '0').split(',')]:
    os.environ['OSTEO_DDPM_DBG'] = str(dbg)
    acc = None
    for rep in range(4):
        n = lib.osteo_ddpm_profile_step(model._ctx, rows, t, 9, 0, buf, 4096, _lib.stream_handle())
        v = [buf[i] for i in range(n)]
        if rep:
            acc = v if acc is None else [a + b for a, b in zip(acc, v)]
    acc = [a / 3 for a in acc]
    k = len(acc) // 12 if len(acc) >= 12 else 1
    ddpm = sum(acc[11::12]); inp = sum(acc[0::12]); hid = sum(acc) - ddpm - inp
    print(f"dbg={dbg:2d} t={t}: step {sum(acc):.3f} ms  input_proj {inp:.3f}  hidden {hid:.3f}  output_proj+update {ddpm:.3f}", flush=True)
