"""ORACLE support — stage the unmodified reference sources so they travel to the GPU box.

    python -m oracle.stage_reference

Copies the reference's Python sources and config (models/, utils/, data/, config/, main.py) from /root/reference into
oracle/_ref/reference/ — a build OUTPUT like the compiled .so: git-ignored, never committed, not gpurun-ignored.
__graft_entry__.build() runs this in the build container; on the GPU box only the staged copy exists.
Consumers (tests/, bench.py --impl reference, scripts/) reach it through oracle/reference_import.py; the product
package never imports it.
"""
from __future__ import annotations

import shutil
import sys
from pathlib import Path

SOURCE = Path("/root/reference")
DEST = Path(__file__).resolve().parent / "_ref" / "reference"
PARTS = ["models", "utils", "data", "config", "main.py"]


def stage(force: bool = False) -> Path | None:
    """Returns the staged root, or None when neither the source tree nor an earlier staging exists."""
    if not SOURCE.is_dir():
        return DEST if (DEST / "models" / "diffusion.py").is_file() else None
    for part in PARTS:
        src, dst = SOURCE / part, DEST / part
        if not src.exists():
            continue
        if src.is_dir():
            for f in src.rglob("*"):
                if f.is_file() and f.suffix in (".py", ".yaml", ".yml") and "__pycache__" not in f.parts:
                    out = dst / f.relative_to(src)
                    if force or not out.exists() or out.read_bytes() != f.read_bytes():
                        out.parent.mkdir(parents=True, exist_ok=True)
                        shutil.copyfile(f, out)
        else:
            dst.parent.mkdir(parents=True, exist_ok=True)
            if force or not dst.exists() or dst.read_bytes() != src.read_bytes():
                shutil.copyfile(src, dst)
    return DEST


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
