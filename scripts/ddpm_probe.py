"""Diagnostic: per-kernel times of one reverse step at t=500 under the OSTEO_DDPM_DBG switches (comma-separated list in PROBE_DBG).
Fused step (fused_step.cuh) bits: 1 = no L2 prefetch of the state, 4 = no state store, 8 = no noise (sigma = 0), 16 = no state load,
64 / 128 = skip the eps / next-input_proj MMAs (timing probes only: results are wrong)."""
import os, sys, ctypes as C
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
if os.environ.get("PROBE_TRACE"):
    os.environ["OSTEO_DDPM_TRACE"] = "1"      # event timeline of CTA 0 on stderr (not legal under graph capture: set after the warm-up)
fused = bool(lib.osteo_ddpm_step_is_fused(model._ctx))
per_chunk = 11 if fused else 12
buf = (C.c_float * 4096)()
for t, dbg in [(500, d) for d in os.environ.get('PROBE_DBG', '0').split(',')]:
    # an entry is 'dbg' or 'dbg:pf' (pf = L2 prefetch distance of the fused kernel, OSTEO_FUSED_PF)
    dbg, _, pf = dbg.partition(':')
    dbg = int(dbg)
    os.environ['OSTEO_DDPM_DBG'] = str(dbg)
    if pf:
        os.environ['OSTEO_FUSED_PF'] = pf
    else:
        os.environ.pop('OSTEO_FUSED_PF', None)
    acc = None
    for rep in range(4):
        n = lib.osteo_ddpm_profile_step(model._ctx, rows, t, 9, 0, buf, 4096, _lib.stream_handle())
        v = [buf[i] for i in range(n)]
        if rep:
            acc = v if acc is None else [a + b for a, b in zip(acc, v)]
    acc = [a / 3 for a in acc]
    ddpm = sum(acc[per_chunk - 1::per_chunk])
    inp = 0.0 if fused else sum(acc[0::per_chunk])
    hid = sum(acc) - ddpm - inp
    if os.environ.get("PROBE_KERNELS"):
        print("  per launch (ms): " + " ".join(f"{a:.4f}" for a in acc[:per_chunk]), flush=True)
    print(f"dbg={dbg:2d} pf={pf or '-'} t={t} fused={int(fused)}: step {sum(acc):.3f} ms  input_proj {inp:.3f}  hidden {hid:.3f}  output_proj+update {ddpm:.3f}", flush=True)
