"""BASELINE.json configs[2] AT SIZE: batch-sharded sampling of 10 M patients across the GPUs of one box, no per-step collective, with the
generation egress (threshold / bit-pack / column split on the device, pinned double buffer, .npy shards + manifest per rank) INSIDE the
timed region. One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 scripts/run_config2_sampling.py [--patients 10000000] [--out /dev/shm/osteo_cfg2]

Wall-clock (max over ranks) for the whole job including file writes; also the same cohort sampled WITHOUT egress for comparison.
If the output directory cannot hold the cohort (20.6 KB per patient), the egress leg runs on as many patients as fit and says so."""
import argparse, json, os, shutil, sys, time
sys.path.insert(0, ".")
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--patients", type=int, default=10_000_000)
ap.add_argument("--out", default="/dev/shm/osteo_cfg2")
ap.add_argument("--shard-rows", type=int, default=100_000)
ap.add_argument("--skip-plain", action="store_true")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    saved = os.dup(1); os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier(); torch.cuda.synchronize()
    sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)

from osteosarcoma_diffusionmodel_b200 import distributed as D
from osteosarcoma_diffusionmodel_b200 import synthetic as synth
from osteosarcoma_diffusionmodel_b200.diffusion import BiologyAwareDiffusionModel

model = BiologyAwareDiffusionModel(62, 5054, 26, 3, synth.model_config())
model.load_state_dict(synth.make_params(5142, 3, (256, 512, 256), seed=0), strict=False)
model = model.to(dev).eval()
n = args.patients
b, e = D.shard_rows(n, rank, world)
idx = torch.arange(b, e, device=dev)
table = torch.tensor(synth.SCENARIO_CONDITIONS, device=dev, dtype=torch.float32)
cond_local = table[(3 * idx) // n]                      # the three config.yaml scenarios in equal thirds of the GLOBAL cohort


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def maxf(v):
    t = torch.tensor([v], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()

model.sample(cond_local[:args.shard_rows], min(args.shard_rows, e - b), seed=1, row_base=b, t_stop=990)      # warm-up
out = {"config": "BASELINE.json configs[2]", "patients": n, "n_gpus": world, "rows_per_gpu": e - b, "shard_rows": args.shard_rows}
if not args.skip_plain:
    barrier(); t0 = time.perf_counter()
    for r0 in range(0, e - b, args.shard_rows):
        r1 = min(r0 + args.shard_rows, e - b)
        model.sample(cond_local[r0:r1], r1 - r0, seed=0, row_base=b + r0)
    barrier(); dt = maxf(time.perf_counter() - t0)
    out["sampling_only"] = {"seconds": dt, "patients_per_s": n / dt, "note": "samples left in HBM, shard by shard"}
# egress leg
# 10 M patients are 206 GB -- more than the RAM-backed scratch of a node: every finished shard is handed to a consumer callback that
# (standing in for shipping it off the box) deletes it, so the space in use stays at ~2 shards per rank while every byte is still written.
ap_free = shutil.disk_usage(os.path.dirname(args.out.rstrip("/")) or "/").free
n_eg = n if ap_free > 4 * world * args.shard_rows * 20_600 else 0
out["egress_patients"] = n_eg
written = {"bytes": 0}


def consume(entry, paths):
    for pth in paths:
        written["bytes"] += os.path.getsize(pth)
        os.remove(pth)
if n_eg > 0:
    be, ee = D.shard_rows(n_eg, rank, world)
    idx = torch.arange(be, ee, device=dev)
    cond_e = table[(3 * idx) // n_eg]
    from osteosarcoma_diffusionmodel_b200.egress import generate_to_shards
    barrier(); t0 = time.perf_counter()
    man = generate_to_shards(model, cond_e, os.path.join(args.out, f"rank_{rank:03d}"), shard_rows=args.shard_rows, seed=0, row_base=be, pack_bits=True,
                             on_shard=consume)
    barrier(); dt = maxf(time.perf_counter() - t0)
    t = torch.tensor([float(written["bytes"])], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t)
    out["with_egress"] = {"seconds": dt, "patients_per_s": n_eg / dt, "bytes_written": t.item(), "gb_per_s": t.item() / dt / 1e9, "dir": args.out,
                          "shards_per_rank": len(man["shards"]), "format": ".npy shards: fp32 expression + pathways, bit-packed mutation calls, conditions",
                          "note": "every shard is fully written to the RAM-backed scratch and then consumed (deleted) by the on_shard callback: 206 GB do not fit the node"}
    barrier()
    shutil.rmtree(os.path.join(args.out, f"rank_{rank:03d}"), ignore_errors=True)
model.check_status()
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
