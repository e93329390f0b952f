"""ORACLE support — import the unmodified reference from /root/reference (build container only).

`models/diffusion.py:9` imports torch_geometric (used only by the never-instantiated
PathwayGraphEncoder, models/diffusion.py:14-88); a two-symbol stub lets the module import.
/root/reference does not exist on the GPU box; what travels there is the staged copy oracle/_ref/reference
(oracle/stage_reference.py, run by __graft_entry__.build(); git-ignored like the compiled .so). This module resolves
whichever exists. Only tests/, bench.py's reference arm and scripts/ call it -- never the product package.
"""
from __future__ import annotations

import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference")


def _resolve_root() -> str:
    env = os.environ.get("OSTEO_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _STAGED):
        if os.path.isfile(os.path.join(cand, "models", "diffusion.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _resolve_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "diffusion.py"))


def _install_stub() -> None:
    if "torch_geometric" in sys.modules:
        return
    tg = types.ModuleType("torch_geometric")
    nn = types.ModuleType("torch_geometric.nn")

    class GATConv:  # pragma: no cover - dead symbol
        def __init__(self, *a, **k):
            raise RuntimeError("torch_geometric stub: GATConv is not on the hot path")

    def global_mean_pool(*a, **k):  # pragma: no cover
        raise RuntimeError("torch_geometric stub")

    nn.GATConv = GATConv
    nn.global_mean_pool = global_mean_pool
    tg.nn = nn
    sys.modules["torch_geometric"] = tg
    sys.modules["torch_geometric.nn"] = nn


def import_reference():
    """Returns (models.diffusion module, utils.validation module) of the reference."""
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_stub()
    import importlib.util

    def load(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    return load("_ref_models_diffusion", "models/diffusion.py"), load("_ref_utils_validation", "utils/validation.py")


class dropin_path:
    """Context manager: sys.path as a reference user would set it to swap in the B200-native classes -- the repo root FIRST (its
    models/diffusion.py and utils/validation.py shims), the reference tree second (everything else: utils/train.py, utils/generate.py,
    models/cvae.py ...). Both trees use namespace packages (no __init__.py, like the reference: main.py:14), so `models` and `utils`
    merge. Modules imported inside are dropped again on exit."""

    def __init__(self, repo_root: str):
        self.repo_root = repo_root

    def __enter__(self):
        if not available():
            raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
        _install_stub()
        self._saved_path = list(sys.path)
        self._saved_mods = {k: v for k, v in sys.modules.items() if k in ("models", "utils", "main") or k.startswith(("models.", "utils."))}
        for k in self._saved_mods:
            del sys.modules[k]
        sys.path[:] = [self.repo_root, REFERENCE_ROOT] + [p for p in sys.path if p not in (self.repo_root, REFERENCE_ROOT, "")]
        import importlib

        importlib.invalidate_caches()
        return self

    def __exit__(self, *exc):
        for k in [k for k in sys.modules if k in ("models", "utils", "main") or k.startswith(("models.", "utils."))]:
            del sys.modules[k]
        sys.modules.update(self._saved_mods)
        sys.path[:] = self._saved_path
