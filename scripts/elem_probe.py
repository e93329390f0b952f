"""Diagnostic: achieved HBM bandwidth of the standalone elementwise kernels of the path (q_sample, reverse_update, load/store_state,
init_noise, train_prepare via one training forward) on 100k x 5142 fp32, against MEASURED_PEAKS.json."""
import json, sys
sys.path.insert(0, ".")
import torch
from osteosarcoma_diffusionmodel_b200 import _lib, synthetic as synth
from osteosarcoma_diffusionmodel_b200.diffusion import BiologyAwareDiffusionModel

n, D = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000, 5142
try:
    peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbps"]
except Exception:
    peak = 6452.5
model = BiologyAwareDiffusionModel(62, 5054, 26, 3, synth.model_config())
model.load_state_dict(synth.make_params(D, 3, (256, 512, 256), seed=0), strict=False)
model = model.to("cuda").eval()
x0 = torch.randn(n, D, device="cuda")
t = torch.randint(0, 1000, (n,), device="cuda")
noise = torch.randn(n, D, device="cuda")
cond = synth.scenario_conditions(n, 3).cuda()
lib, s = _lib.load(), _lib.stream_handle()
model._ensure_ctx(n)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


row = D * 4
out = torch.empty_like(x0)
eps = torch.randn(n, D, device="cuda")
xs = x0.clone()
cases = {
    "reverse_update (injected z: read x + eps + z, write x)": (lambda: _lib.check(lib.osteo_ddpm_reverse_update(model._ctx, xs.data_ptr(), eps.data_ptr(), noise.data_ptr(), n, 500, 0, 0, s)), 4 * row),
    "reverse_update (Philox z: read x + eps, write x)": (lambda: _lib.check(lib.osteo_ddpm_reverse_update(model._ctx, xs.data_ptr(), eps.data_ptr(), None, n, 500, 7, 0, s)), 3 * row),
    "q_sample (injected noise: read x0 + noise, write x_t)": (lambda: model.q_sample(x0, t, noise), 3 * row),
    "q_sample (Philox noise: read x0, write noise + x_t)": (lambda: model.q_sample(x0, t), 3 * row),
    "load_state (read x, write fp32 state + bf16 shadow)": (lambda: _lib.check(lib.osteo_ddpm_load_state(model._ctx, x0.data_ptr(), n, s)), 2 * row + D * 2),
    "store_state (read state, write x)": (lambda: _lib.check(lib.osteo_ddpm_store_state(model._ctx, out.data_ptr(), n, s)), 2 * row),
    "init_noise (write fp32 state + bf16 shadow)": (lambda: _lib.check(lib.osteo_ddpm_init_noise(model._ctx, n, 1, 0, s)), row + D * 2),
}
for name, (fn, bytes_per_row) in cases.items():
    ms = timed(fn)
    gbs = bytes_per_row * n / (ms / 1e3) / 1e9
    print(f"{name}: {ms:.3f} ms, {gbs:.0f} GB/s algorithmic = {100 * gbs / peak:.0f} % of {peak:.0f} GB/s", flush=True)
