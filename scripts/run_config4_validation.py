"""BASELINE.json configs[4] AT SIZE: RBF-MMD and pathway coherence on 1 M synthetic vs 1 M reference-shaped patients (5142 features),
Gram rows sharded block-cyclically over the GPUs of one box (symmetric half-Grams, NCCL all-reduce of 3 fp64), cohort rows sharded for
the coherence moments. One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/run_config4_validation.py [--rows 1000000]

Every rank generates the same X, Y on its device (same seed), as the validator expects the whole cohorts on every rank. Device-timed
between barriers, max over ranks; rank 0 prints one JSON line."""
import argparse, json, os, sys, time
sys.path.insert(0, ".")
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--features", type=int, default=5142)
ap.add_argument("--precision", default="bf16")
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    saved = os.dup(1); os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier(); torch.cuda.synchronize()
    sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)

from osteosarcoma_diffusionmodel_b200.validation import BiologicalValidator

n, d = args.rows, args.features
g = torch.Generator(device=dev).manual_seed(1)
X = torch.empty((n, d), device=dev); Y = torch.empty((n, d), device=dev)
step = 100_000
for r0 in range(0, n, step):          # chunked generation keeps the temporary small
    X[r0:r0 + step].normal_(generator=g).add_(4.0)
for r0 in range(0, n, step):
    Y[r0:r0 + step].normal_(generator=g).mul_(1.1).add_(4.1)
val = BiologicalValidator({"evaluation": {}}, precision=args.precision)


def timed(fn):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return out, t.item()

val.compute_mmd(X[:8192], Y[:8192])          # warm-up (workspace allocation, function attributes)
mmd, ms = timed(lambda: val.compute_mmd(X, Y))
peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}
tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
passes = 3 if args.precision == "fp32x3" else 1
pairs = 3.0 * n * n
mults = (2 * (n * (n + 128.0) / 2) + n * float(n)) * 2.0 * d * passes            # symmetric halves of Kxx, Kyy + the whole Kxy
out = {"config": "BASELINE.json configs[4]", "rows": n, "features": d, "n_gpus": world, "precision": args.precision,
       "mmd": {"value": mmd, "seconds": ms / 1e3, "kernel_pairs_per_s": pairs / (ms / 1e3), "tensor_tflops_per_gpu": mults / (ms / 1e3) / 1e12 / world,
               "tensor_frac_of_sustained_peak": mults / (ms / 1e3) / 1e12 / world / tf_peak,
               "sharding": "Gram rows block-cyclic over ranks, Kxx / Kyy as symmetric half-Grams, one all-reduce of 3 fp64"}}
del Y
# pathway coherence: 10 pathways x 15 genes out of the first 371 columns of both cohorts (rows sharded, moment blocks all-reduced)
real, syn = X[:, :371].contiguous(), X[:, 371:742].contiguous()
members = [list(range(15 * p, 15 * p + 15)) for p in range(10)]
val2 = BiologicalValidator({"evaluation": {}})
val2.pathway_coherence_from_tensors(real[:4096], syn[:4096], members)
coh, ms = timed(lambda: val2.pathway_coherence_from_tensors(real, syn, members))
hbm = float(peaks.get("hbm_gbs", 6650.0))
out["coherence"] = {"rows_per_cohort": n, "genes": 371, "pathways": 10, "seconds": ms / 1e3, "values": coh,
                    "streamed_gb_per_s_per_gpu": 2.0 * n * 371 * 4 / (ms / 1e3) / 1e9 / world, "hbm_frac_per_gpu": 2.0 * n * 371 * 4 / (ms / 1e3) / 1e9 / world / hbm}
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
