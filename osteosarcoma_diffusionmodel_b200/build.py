"""In-tree build of the CUDA extension: one nvcc invocation, sm_100a only.

    python -m osteosarcoma_diffusionmodel_b200.build [--force]

Produces osteosarcoma_diffusionmodel_b200/libosteo_ddpm.so next to the sources, so the built
library travels with a snapshot of the repo. nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
OUT = HERE / "libosteo_ddpm.so"
STAMP = HERE / ".libosteo_ddpm.stamp"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",
]


# extra nvcc flags for diagnostics builds, e.g. OSTEO_NVCC_EXTRA="-DOSTEO_FUSED_TRACE"
NVCC_FLAGS += os.environ.get("OSTEO_NVCC_EXTRA", "").split()


def _sources():
    files = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.inl"))
    files.append(HERE.parent / "include" / "osteo_ddpm.h")
    return files


def _digest() -> str:
    h = hashlib.sha256()
    for f in _sources():
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: cannot build the CUDA extension (there is no CPU fallback)")


def build(force: bool = False, verbose: bool = False) -> Path:
    digest = _digest()
    if not force and OUT.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return OUT
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", str(OUT), str(CSRC / "osteo_ddpm.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed with exit code {res.returncode}")
    STAMP.write_text(digest)
    return OUT


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv)
    print(p)
