"""Generation egress (SURVEY.md §8f "next" #1): the step after the sampling loop.

The reference moves the whole cohort to the host (`.cpu().numpy()`), splits the columns, thresholds the mutation block and writes four
5142-column CSVs per scenario through pandas (utils/generate.py:127-144, :177-235).  At the 10 M patients of BASELINE.json configs[2]
that is 206 GB of fp32 and the writer, not the GPU, is the wall.  Here the cohort is sampled in shards; the split / threshold / bit
packing happens on the device (`sample_components` -> `osteo_ddpm_store_split`), each shard goes to pinned host buffers on a copy
stream while the next shard is already sampling, and a writer thread stores it as `.npy` files plus a JSON manifest.  `load_shards`
gives back the dictionary `SyntheticPatientGenerator.generate` returns (mutations as 0.0 / 1.0 floats, expression, pathways,
conditions), so the reference's own `save_synthetic_data` can still write CSVs from it at small N.
"""
from __future__ import annotations

import json
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Dict, List, Optional

import numpy as np
import torch


def generate_to_shards(model, conditions: torch.Tensor, out_dir, shard_rows: int = 100_000, seed: int = 0, row_base: int = 0,
                       pack_bits: bool = True, on_shard=None) -> Dict:
    """Sample `conditions.shape[0]` patients shard by shard and write them under `out_dir`; returns the manifest.
    Rows keep their global Philox identity (row_base + index), so the files do not depend on the shard size or on how a cohort was
    split over GPUs (each rank calls this with its own row range and directory).
    on_shard(entry, paths), if given, runs on the writer thread right after a shard's files are complete (e.g. to ship them off the box
    and delete them: 10 M patients are 206 GB, more than a node's RAM-backed scratch)."""
    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    n = int(conditions.shape[0])
    on_gpu = conditions.device.type == "cuda"
    copy_stream = torch.cuda.Stream(device=conditions.device) if on_gpu else None
    manifest = {"format": "osteo-ddpm-b200 shards v1", "rows": n, "shard_rows": int(shard_rows), "seed": int(seed), "row_base": int(row_base),
                "mutation_dim": int(model.mutation_dim), "expression_dim": int(model.expression_dim), "pathway_dim": int(model.pathway_dim),
                "mutations": "bits (LSB first, ceil(mutation_dim / 8) bytes per patient)" if pack_bits else "uint8 0/1", "shards": []}
    pool = ThreadPoolExecutor(max_workers=1)      # the writer: at most one shard in flight on the host
    pending = None                                # (future, manifest entry) of the shard being written
    pinned: List[Dict[str, torch.Tensor]] = [{}, {}]      # two sets of pinned staging buffers, reused shard after shard

    def write(idx: int, host: Dict[str, torch.Tensor], ready) -> None:
        if ready is not None:
            ready.synchronize()
        paths = []
        for k, t in host.items():
            paths.append(out / f"shard_{idx:05d}_{k}.npy")
            np.save(paths[-1], t.numpy())
        return paths

    def finish(p) -> None:
        """Wait for a shard's files; a failure in the writer thread (full disk, permissions, a CUDA error surfacing in
        ready.synchronize()) is re-raised HERE, and only a completely written shard enters the manifest."""
        fut, entry = p
        paths = fut.result()
        if on_shard is not None:
            on_shard(entry, paths)
        manifest["shards"].append(entry)

    def staging(slot: int, key: str, like: torch.Tensor) -> torch.Tensor:
        buf = pinned[slot].get(key)
        if buf is None or buf.dtype != like.dtype or buf.shape[1:] != like.shape[1:] or buf.shape[0] < like.shape[0]:
            buf = torch.empty((max(like.shape[0], min(shard_rows, n)),) + tuple(like.shape[1:]), dtype=like.dtype, pin_memory=True)
            pinned[slot][key] = buf
        return buf[:like.shape[0]]

    try:
        for idx, b in enumerate(range(0, n, shard_rows)):
            e = min(b + shard_rows, n)
            comp = model.sample_components(conditions[b:e], e - b, seed=seed, row_base=row_base + b, pack_bits=pack_bits)
            keep = {"expression": comp["expression"], "pathways": comp["pathways"], "conditions": comp["conditions"],
                    ("mutation_bits" if pack_bits else "mutations"): comp["mutation_bits" if pack_bits else "mutations"]}
            host, ready = {}, None
            if on_gpu:
                copy_stream.wait_stream(torch.cuda.current_stream(conditions.device))
                with torch.cuda.stream(copy_stream):
                    for k, t in keep.items():
                        # the copy reads the SOURCE storage (expression / pathways are column views of one buffer) on copy_stream: tell the
                        # caching allocator, or the storage could be handed out again while the copy is still in flight
                        t.record_stream(copy_stream)
                        # slot idx % 2 was last used by shard idx - 2, whose writer finished before shard idx - 1's writer started
                        host[k] = staging(idx % 2, k, t)
                        host[k].copy_(t, non_blocking=True)          # strided device view -> dense pinned rows in one 2-D copy
                    ready = torch.cuda.Event()
                    ready.record(copy_stream)
            else:
                host = {k: t.contiguous() for k, t in keep.items()}
            if pending is not None:
                finish(pending)
            pending = (pool.submit(write, idx, host, ready), {"index": idx, "row_begin": row_base + b, "rows": e - b})
        if pending is not None:
            finish(pending)
            pending = None
    finally:
        pool.shutdown(wait=True)
    (out / "manifest.json").write_text(json.dumps(manifest, indent=1))
    return manifest


def load_shards(out_dir, shards: Optional[List[int]] = None) -> Dict[str, np.ndarray]:
    """The dictionary of SyntheticPatientGenerator.generate (utils/generate.py:137-144) from a shard directory: 'mutations' as 0.0 / 1.0
    float64 (the reference's `(mutations > 0.5).astype(float)`), 'expression', 'pathways', 'conditions'."""
    out = Path(out_dir)
    man = json.loads((out / "manifest.json").read_text())
    md = man["mutation_dim"]
    parts: Dict[str, list] = {"mutations": [], "expression": [], "pathways": [], "conditions": []}
    for sh in man["shards"]:
        if shards is not None and sh["index"] not in shards:
            continue
        i = sh["index"]
        if (out / f"shard_{i:05d}_mutation_bits.npy").exists():
            bits = np.load(out / f"shard_{i:05d}_mutation_bits.npy")
            mut = np.unpackbits(bits, axis=1, bitorder="little")[:, :md]
        else:
            mut = np.load(out / f"shard_{i:05d}_mutations.npy")
        parts["mutations"].append(mut.astype(float))
        for k in ("expression", "pathways", "conditions"):
            parts[k].append(np.load(out / f"shard_{i:05d}_{k}.npy"))
    return {k: np.concatenate(v) if v else np.zeros((0,)) for k, v in parts.items()}
