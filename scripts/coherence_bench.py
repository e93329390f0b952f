import sys, time
sys.path.insert(0, ".")
import numpy as np, torch
from osteosarcoma_diffusionmodel_b200 import _lib
from osteosarcoma_diffusionmodel_b200.validation import BiologicalValidator
lib = _lib.load(); s = _lib.stream_handle()
rows, genes = 500_000, 371
g = torch.Generator(device="cuda").manual_seed(0)
data = torch.randn(rows, genes, device="cuda", generator=g)
rng = np.random.RandomState(0)
members = [sorted(rng.choice(genes, 15, replace=False).tolist()) for _ in range(10)]
ci = np.full((10, 32), -1, np.int32)
for i, c in enumerate(members): ci[i, :15] = c
ci_t = torch.from_numpy(ci).cuda()
shift = data[0, ci_t.clamp(min=0).long()].contiguous()
out = torch.empty((10, 1057), dtype=torch.float64, device="cuda")
def run():
    _lib.check(lib.osteo_corr_moments_batched(data.data_ptr(), rows, genes, genes, ci_t.data_ptr(), 10, shift.data_ptr(), 0, rows, out.data_ptr(), s))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
print("batched kernel ms per 500k-row cohort:", e0.elapsed_time(e1) / 10)
val = BiologicalValidator({}, device="cuda")
t0 = time.perf_counter()
for _ in range(5): val._coherence_scores(data, members)
torch.cuda.synchronize()
print("python _coherence_scores ms:", (time.perf_counter() - t0) / 5 * 1e3)
