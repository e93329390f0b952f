"""CPU: the C-ABI library builds, loads and exports exactly the symbols include/osteo_ddpm.h declares."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "osteo_ddpm.h"


def header_symbols():
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(osteo_[a-z0-9_]+)\s*\(", text))


@pytest.fixture(scope="module")
def lib():
    from osteosarcoma_diffusionmodel_b200 import build, _lib

    build.build()
    return _lib.load()


def test_header_and_binding_declare_the_same_symbols(lib):
    from osteosarcoma_diffusionmodel_b200 import _lib

    hs = header_symbols()
    assert hs == set(_lib.SIGNATURES), (hs ^ set(_lib.SIGNATURES))
    assert len(hs) >= 30


def test_library_exports_every_declared_symbol(lib):
    for name in header_symbols():
        assert hasattr(lib, name), name


def test_version_and_weight_count(lib):
    assert lib.osteo_version() >= 100
    assert lib.osteo_ddpm_num_weight_tensors(3) == 52      # 52 parameter tensors for hidden [256,512,256]
    assert lib.osteo_ddpm_num_weight_tensors(2) == 36


def test_no_cpu_fallback(lib):
    """Without a CUDA device every compute entry point fails loudly instead of computing on the host."""
    import ctypes as C
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    assert lib.osteo_device_count() == 0
    h = C.c_void_p()
    hid = (C.c_int * 3)(256, 512, 256)
    assert lib.osteo_ddpm_create(C.byref(h), 0, 5142, 3, 128, 64, 3, hid, 1000, 0.2, 0) != 0
    assert b"no CUDA device" in lib.osteo_last_error()
    assert lib.osteo_philox_normal(None, 4, 4, 0, 0, 0, 0, None) != 0


def test_model_refuses_to_compute_on_cpu():
    import torch
    from oracle import synth
    from osteosarcoma_diffusionmodel_b200.diffusion import BiologyAwareDiffusionModel

    m = BiologyAwareDiffusionModel(100, 200, 50, 5, synth.model_config())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m.sample(torch.zeros(2, 5), num_samples=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(2, 350), torch.zeros(2, 5))


def test_product_package_never_imports_the_oracle():
    pkg = ROOT / "osteosarcoma_diffusionmodel_b200"
    for f in list(pkg.rglob("*.py")) + [ROOT / "models" / "diffusion.py"]:
        txt = f.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
