"""Diagnostic: kernel-only and whole-call times of the pathway-coherence path (1 M rows x 371 genes, 10 pathways x 15 genes)."""
import sys
sys.path.insert(0, ".")
import torch
from osteosarcoma_diffusionmodel_b200 import validation as val
from osteosarcoma_diffusionmodel_b200.validation import BiologicalValidator

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 15
g = torch.Generator(device="cuda").manual_seed(1)
cohort = torch.randn(rows, 371, device="cuda", generator=g)
members = [list(range(k * p, k * p + k)) for p in range(10)]
v = BiologicalValidator({"evaluation": {}})
half = rows // 2
v.pathway_coherence_from_tensors(cohort[:half], cohort[half:], members)
ci_t = v._index_tensor(cohort.device, members)


def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

ms_k = timed(lambda: val._moments_tiled(cohort[:half], ci_t, (0, half), k))
ms_old = timed(lambda: val._moments_batched(cohort[:half], ci_t, cohort[0, ci_t.clamp(min=0).long()].contiguous(), (0, half)))
ms_call = timed(lambda: v.pathway_coherence_from_tensors(cohort[:half], cohort[half:], members))
gb = half * 371 * 4 / 1e9
print(f"rows/cohort {half} k={k}: tiled kernel {ms_k:.3f} ms ({gb / ms_k * 1e3:.0f} GB/s), warp-per-set kernel {ms_old:.3f} ms, whole call (2 cohorts + finish + D2H) {ms_call:.3f} ms ({2 * gb / ms_call * 1e3:.0f} GB/s)")
