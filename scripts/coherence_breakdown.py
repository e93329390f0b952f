"""Where the public pathway-coherence call spends its time at the bench size (1 M rows x 371 genes, two cohorts of 500k):
device time of the two moment kernels + finish (CUDA events, no host sync inside) vs the whole call (host work + D2H sync)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import torch
from osteosarcoma_diffusionmodel_b200.validation import BiologicalValidator, _coherence_finish, _CM_STRIDE

dev = torch.device("cuda")
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
cohort = torch.randn(rows, 371, device=dev)
members = [list(range(15 * p, 15 * p + 15)) for p in range(10)]
val = BiologicalValidator({"evaluation": {}})
a, b = cohort[: rows // 2], cohort[rows // 2:]
for _ in range(2):
    val.pathway_coherence_from_tensors(a, b, members)
torch.cuda.synchronize()

def ev(fn, reps=5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

mom = torch.empty((20, _CM_STRIDE), dtype=torch.float64, device=dev)
ci = torch.cat([val._index_tensor(dev, members), val._index_tensor(dev, members)])
def device_only():
    val._coherence_moments(a, members, out=mom[:10], reduce=False)
    val._coherence_moments(b, members, out=mom[10:], reduce=False)
    _coherence_finish(mom, ci)
print(f"device work only (2 moment kernels + finish, no sync): {ev(device_only):.3f} ms")
print(f"whole public call: {ev(lambda: val.pathway_coherence_from_tensors(a, b, members)):.3f} ms")
t0 = time.perf_counter()
for _ in range(20):
    val.pathway_coherence_from_tensors(a, b, members)
torch.cuda.synchronize()
print(f"whole public call, wall clock: {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms")
x = np.random.rand(10); y = np.random.rand(10)
t0 = time.perf_counter()
for _ in range(100):
    np.corrcoef(x, y)[0, 1]; np.mean(x); np.mean(y)
print(f"numpy tail (corrcoef + 2 means): {(time.perf_counter() - t0) / 100 * 1e3:.3f} ms")
