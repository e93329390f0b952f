"""Multi-GPU plumbing for the hot path: one process per GPU, torch.distributed (NCCL over
NVLink on GPUs, gloo in the CPU tests). The data path itself never communicates:

  sampling   rows are independent -> contiguous global-row shards, no collective (Philox is keyed by global row);
  training   data parallel -> ONE all-reduce (mean) of the model's flat gradient buffer (generic modules: bucketed all-reduce of the
             .grad tensors, largest / last-layer bucket first);
  MMD        Gram ROWS sharded, one all-reduce of the three fp64 partial sums;
  coherence  cohort rows sharded, one all-reduce of the per-pathway fp64 moment blocks.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import torch

try:  # torch.distributed is always present in this image; keep the import soft for documentation builds
    import torch.distributed as dist
except Exception:  # pragma: no cover
    dist = None


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when no process group is initialised."""
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_rows(n: int, rank: int, world_size: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous row range [begin, end) of rank `rank`; every boundary except n is a multiple of `align`.
    Ranges tile [0, n) exactly, trailing ranks may be empty."""
    if n < 0 or world_size < 1 or not (0 <= rank < world_size) or align < 1:
        raise ValueError("bad shard arguments")
    units = (n + align - 1) // align
    per = (units + world_size - 1) // world_size
    begin = min(n, rank * per * align)
    end = min(n, (rank + 1) * per * align)
    return begin, end


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    rank, ws = world()
    if ws > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def make_buckets(tensors: Sequence[torch.Tensor], bucket_bytes: int) -> List[List[int]]:
    """Indices of `tensors` grouped in REVERSE order (the backward pass finishes output_proj first) into buckets of at
    most `bucket_bytes` (a tensor larger than the cap gets its own bucket)."""
    buckets, cur, size = [], [], 0
    for i in reversed(range(len(tensors))):
        b = tensors[i].numel() * tensors[i].element_size()
        if cur and size + b > bucket_bytes:
            buckets.append(cur)
            cur, size = [], 0
        cur.append(i)
        size += b
    if cur:
        buckets.append(cur)
    return buckets


def allreduce_gradients(params: Iterable[torch.nn.Parameter], bucket_bytes: int = 8 << 20) -> int:
    """Average .grad over the data-parallel group with one flat all-reduce per bucket, all launched asynchronously and then
    waited in order (NCCL serialises them on its stream; buckets keep the 17 MB payload latency-bound pieces few).
    Returns the number of buckets (0 when world_size == 1)."""
    rank, ws = world()
    ps = [p for p in params if p.grad is not None]
    if ws == 1 or not ps:
        return 0
    grads = [p.grad for p in ps]
    buckets = make_buckets(grads, bucket_bytes)
    flats, works = [], []
    for b in buckets:
        flat = torch.cat([grads[i].reshape(-1) for i in b])
        works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True))
        flats.append(flat)
    for b, flat, w in zip(buckets, flats, works):
        w.wait()
        flat.div_(ws)
        o = 0
        for i in b:
            n = grads[i].numel()
            grads[i].copy_(flat[o:o + n].view_as(grads[i]))
            o += n
    return len(buckets)


def dp_train_step(model, optimizer, x0: torch.Tensor, conditions: torch.Tensor, max_grad_norm: float = 1.0, bucket_bytes: int = 8 << 20) -> torch.Tensor:
    """One data-parallel optimiser step with the reference's recipe (utils/train.py:230-246): the global gradient norm is
    taken AFTER the all-reduce, so every rank clips identically and parameters stay bit-identical across ranks."""
    optimizer.zero_grad()
    flat = hasattr(model, "_flat_allreduce") and hasattr(model, "_grad_buf")      # BiologyAwareDiffusionModel: one all-reduce of its flat buffer
    if flat:
        model._flat_allreduce = True
    try:
        loss = model(x0, conditions, return_loss=True)
        loss.backward()
    finally:
        if flat:
            model._flat_allreduce = False
    if not flat:
        allreduce_gradients(model.parameters(), bucket_bytes)
    # optim.FusedAdamW(max_grad_norm=...) clips inside its own two-launch step (global norm on the device, after the reduce): no torch clip then
    if getattr(optimizer, "defaults", {}).get("max_grad_norm") is None:
        torch.nn.utils.clip_grad_norm_(model.parameters(), max_grad_norm)
    optimizer.step()
    return loss.detach()


def sample_sharded(model, conditions: torch.Tensor, num_samples: int, seed: int = 0, gather: bool = False) -> torch.Tensor:
    """Batch-sharded sampling (BASELINE.json configs[2]): `conditions` holds the GLOBAL cohort's condition rows; every rank
    generates its contiguous slice with row_base = first global row, so the union is identical for any GPU count.
    Returns the local slice, or the full cohort on every rank when gather=True."""
    rank, ws = world()
    b, e = shard_rows(num_samples, rank, ws)
    local = model.sample(conditions[b:e], e - b, seed=seed, row_base=b) if e > b else torch.empty((0, model.data_dim), device=conditions.device)
    if not gather or ws == 1:
        return local
    sizes = [shard_rows(num_samples, r, ws) for r in range(ws)]
    outs = [torch.empty((hi - lo, model.data_dim), device=local.device, dtype=local.dtype) for lo, hi in sizes]
    dist.all_gather(outs, local) if len({o.shape for o in outs}) == 1 else _all_gather_ragged(outs, local, sizes)
    return torch.cat(outs)


def sample_sharded_to_shards(model, conditions: torch.Tensor, num_samples: int, out_dir, seed: int = 0, rows_per_shard: int = 100_000, pack_bits: bool = True):
    """Batch-sharded sampling straight to disk (BASELINE.json configs[2]: 10 M patients): rank r samples its contiguous slice of the global
    cohort and streams it to `out_dir/rank_{r:03d}` with egress.generate_to_shards (rows keep their global index, so the union of the
    directories is the same cohort for any GPU count). No collective. Returns this rank's manifest."""
    from pathlib import Path

    from .egress import generate_to_shards
    rank, ws = world()
    b, e = shard_rows(num_samples, rank, ws)
    return generate_to_shards(model, conditions[b:e], Path(out_dir) / f"rank_{rank:03d}", shard_rows=rows_per_shard, seed=seed, row_base=b, pack_bits=pack_bits)


def _all_gather_ragged(outs, local, sizes):
    for r, o in enumerate(outs):
        if o.numel():
            dist.broadcast(o if r != dist.get_rank() else local, src=r)
            if r == dist.get_rank():
                o.copy_(local)
