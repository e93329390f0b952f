"""ORACLE support — import the unmodified reference from /root/reference (build container only).

`models/diffusion.py:9` imports torch_geometric (used only by the never-instantiated
PathwayGraphEncoder, models/diffusion.py:14-88); a two-symbol stub lets the module import.
/root/reference does not exist on the GPU box: nothing under tests -m gpu, smoke() or bench.py
calls this module.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("OSTEO_REFERENCE_ROOT", "/root/reference")


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
