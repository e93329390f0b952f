"""ctypes binding of the C-ABI in include/osteo_ddpm.h.

The shared library is built in-tree by ``build.py`` (nvcc, sm_100a). There is no CPU
fallback: a missing library or a missing CUDA device raises, it never degrades silently.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libosteo_ddpm.so"

PREC_BF16 = 0
PREC_FP32X3 = 1

_c_f32p = C.c_void_p
_vp = C.c_void_p
_ll = C.c_longlong
_u64 = C.c_uint64
_u32 = C.c_uint32
_i = C.c_int
_f = C.c_float
_d = C.c_double

# name -> (restype, argtypes). Mirrors include/osteo_ddpm.h one to one; tests/test_abi.py checks the
# header and this table declare the same symbols.
SIGNATURES = {
    "osteo_last_error": (C.c_char_p, []),
    "osteo_version": (_i, []),
    "osteo_device_count": (_i, []),
    "osteo_ddpm_num_weight_tensors": (_i, [_i]),
    "osteo_ddpm_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, _i, C.POINTER(_i), _i, _f, _i]),
    "osteo_ddpm_destroy": (_i, [_vp]),
    "osteo_ddpm_reserve": (_i, [_vp, _ll]),
    "osteo_ddpm_capacity": (_ll, [_vp]),
    "osteo_ddpm_workspace_bytes": (_ll, [_vp]),
    "osteo_ddpm_set_chunk_rows": (_i, [_vp, _i]),
    "osteo_ddpm_set_precision": (_i, [_vp, _i]),
    "osteo_ddpm_set_fused": (_i, [_vp, _i]),
    "osteo_ddpm_set_branches": (_i, [_vp, _i]),
    "osteo_ddpm_set_train_graph": (_i, [_vp, _i]),
    "osteo_ddpm_step_is_fused": (_i, [_vp]),
    "osteo_ddpm_graph_branches": (_i, [_vp]),
    "osteo_ddpm_set_weights": (_i, [_vp, C.POINTER(_vp), _i, _vp]),
    "osteo_ddpm_set_schedule": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "osteo_ddpm_set_time_embedding": (_i, [_vp, _vp]),
    "osteo_ddpm_load_state": (_i, [_vp, _vp, _ll, _vp]),
    "osteo_ddpm_store_state": (_i, [_vp, _vp, _ll, _vp]),
    "osteo_ddpm_store_split": (_i, [_vp, _ll, _i, _f, _vp, _vp, _vp, _vp]),
    "osteo_ddpm_init_noise": (_i, [_vp, _ll, _u64, _ll, _vp]),
    "osteo_ddpm_set_conditions": (_i, [_vp, _vp, _ll, _vp]),
    "osteo_ddpm_reverse_step": (_i, [_vp, _ll, _i, _vp, _vp, _u64, _ll, _vp]),
    "osteo_ddpm_sample_loop": (_i, [_vp, _ll, _i, _i, _vp, _u64, _ll, _i, _vp]),
    "osteo_ddpm_denoise": (_i, [_vp, _vp, _vp, _ll, _vp, _vp]),
    "osteo_ddpm_q_sample": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _i, _u64, _ll, _u32, _vp]),
    "osteo_ddpm_reverse_update": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _u64, _ll, _vp]),
    "osteo_ddpm_train_step": (_i, [_vp, _vp, _vp, _ll, _vp, _vp, C.POINTER(_vp), _i, _u64, _ll, _vp, C.POINTER(_vp), _i, _vp]),
    "osteo_ddpm_train_forward": (_i, [_vp, _vp, _vp, _ll, _vp, _vp, C.POINTER(_vp), _i, _u64, _ll, _vp, _vp]),
    "osteo_ddpm_train_x0hat": (_i, [_vp, _vp, _vp, _ll, _vp, _i, _vp, _vp]),
    "osteo_ddpm_train_inject": (_i, [_vp, _vp, _ll, _vp, _i, _vp, _vp]),
    "osteo_ddpm_train_backward": (_i, [_vp, _vp, _ll, _vp, C.POINTER(_vp), _i, _u64, _ll, C.POINTER(_vp), _i, _vp]),
    "osteo_ddpm_train_backward_part": (_i, [_vp, _vp, _ll, _vp, C.POINTER(_vp), _i, _u64, _ll, C.POINTER(_vp), _i, _i, _i, _vp]),
    "osteo_ddpm_enable_training": (_i, [_vp, _i]),
    "osteo_adamw_create": (_i, [C.POINTER(_vp), _i, C.POINTER(_ll)]),
    "osteo_adamw_destroy": (_i, [_vp]),
    "osteo_adamw_step": (_i, [_vp, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_vp), _d, _d, _d, _d, _d, _ll, _d, _vp, _vp]),
    "osteo_ddpm_profile_step": (_i, [_vp, _ll, _i, _u64, _ll, _vp, _i, _vp]),
    "osteo_ddpm_status": (_i, [_vp, _vp]),
    "osteo_ddpm_launch_count": (_ll, [_vp]),
    "osteo_linear_tc": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "osteo_linear_gn_silu_tc": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "osteo_wgrad_tc": (_i, [_vp, _vp, _vp, _ll, _i, _i, _i, _vp]),
    "osteo_philox_normal": (_i, [_vp, _ll, _i, _u64, _ll, _u32, _u32, _vp]),
    "osteo_philox_words": (_i, [_vp, _ll, _i, _u64, _ll, _u32, _u32, _vp]),
    "osteo_mmd_partial": (_i, [_vp, _ll, _vp, _ll, _i, _f, _vp, _ll, _ll, _ll, _ll, _i, _vp, _vp]),
    "osteo_mmd_partial_cyclic": (_i, [_vp, _ll, _vp, _ll, _i, _f, _vp, _i, _i, _i, _vp, _vp]),
    "osteo_corr_moments": (_i, [_vp, _ll, _i, _vp, _i, _vp, _ll, _ll, _vp, _vp]),
    "osteo_corr_moments_batched": (_i, [_vp, _ll, _i, _i, _vp, _i, _vp, _ll, _ll, _vp, _vp]),
    "osteo_corr_moments_tiled": (_i, [_vp, _ll, _i, _i, _vp, _i, _i, _vp, _ll, _ll, _vp, _vp]),
    "osteo_coherence_finish": (_i, [_vp, _vp, _i, _vp, _vp]),
    "osteo_mixup_rows": (_i, [_vp, _ll, _i, _vp, _vp, _ll, _f, _f, _vp, _vp]),
    "osteo_corr_loss_finish": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp]),
    "osteo_corr_loss_backward": (_i, [_vp, _ll, _i, _vp, _i, _vp, _vp, _vp, _vp]),
}


class OsteoError(RuntimeError):
    """Raised when a C-ABI call returns non-zero."""


_lib = None


def load() -> C.CDLL:
    """Load the in-tree shared library; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("OSTEO_DDPM_LIB", LIB_PATH))
    if not path.exists():
        raise OsteoError(
            f"{path} not found: build the CUDA extension first (python -m osteosarcoma_diffusionmodel_b200.build). "
            "There is no CPU fallback."
        )
    lib = C.CDLL(str(path))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().osteo_last_error()
        raise OsteoError(msg.decode() if msg else f"osteo call failed with code {rc}")


def ptr(t) -> int | None:
    """Device (or host) address of a torch tensor / numpy array, None passes NULL."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def stream_handle(device=None) -> int:
    """cudaStream_t of torch's current stream on `device` (default: the current device; the callers run under
    `torch.cuda.device(model device)`, see on_device)."""
    import torch

    return torch.cuda.current_stream(device).cuda_stream


def on_device(device_of):
    """Decorator factory for methods that call the C-ABI: run the method with `device_of(self)` as the current CUDA device, so
    allocations, torch's current stream and the library's launches all refer to the object's device -- not to whatever device the
    process happens to have current (a model on cuda:1 in a process whose current device is cuda:0)."""
    import functools

    def deco(fn):
        @functools.wraps(fn)
        def wrapper(self, *args, **kwargs):
            import torch

            dev = device_of(self)
            if dev is not None and getattr(dev, "type", None) == "cuda" and torch.cuda.is_available():
                with torch.cuda.device(dev):
                    return fn(self, *args, **kwargs)
            return fn(self, *args, **kwargs)
        return wrapper
    return deco
