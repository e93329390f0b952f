"""First-light check of the tcgen05 GEMM on a B200 (run under gpurun). Not a pytest file."""
import sys, time
import ctypes as C
sys.path.insert(0, ".")
import torch
from osteosarcoma_diffusionmodel_b200 import _lib

lib = _lib.load()
print("devices", lib.osteo_device_count(), torch.cuda.get_device_name(0), flush=True)
torch.manual_seed(0)
dev = "cuda"

def run_linear(m, n, k, prec, bias=True):
    a = torch.randn(m, k, device=dev)
    w = torch.randn(n, k, device=dev) / k ** 0.5
    b = torch.randn(n, device=dev) if bias else None
    out = torch.full((m, n), float("nan"), device=dev)
    rc = lib.osteo_linear_tc(a.data_ptr(), w.data_ptr(), b.data_ptr() if bias else None, out.data_ptr(), m, n, k, prec, None)
    if rc != 0:
        print(f"linear m={m} n={n} k={k} prec={prec}: ERROR {lib.osteo_last_error().decode()}", flush=True)
        return
    ref = (a.double() @ w.double().t() + (b.double() if bias else 0)).float()
    if prec == 0:
        ref_bf = (a.bfloat16().double() @ w.bfloat16().double().t() + (b.double() if bias else 0)).float()
    else:
        ref_bf = ref
    err = (out - ref).abs().max().item()
    err_bf = (out - ref_bf).abs().max().item()
    rel = ((out - ref).norm() / ref.norm()).item()
    nan = torch.isnan(out).sum().item()
    print(f"linear m={m} n={n} k={k} prec={prec}: max_abs_err_vs_fp64={err:.3e} vs_bf16_inputs={err_bf:.3e} rel_fro={rel:.3e} nan={nan}", flush=True)

for (m, n, k) in [(128, 128, 64), (128, 128, 256), (256, 256, 512), (100, 128, 64), (300, 512, 1024), (1000, 256, 5142), (777, 5142, 256), (4096, 512, 512)]:
    for prec in (0, 1):
        run_linear(m, n, k, prec)

def run_gn(m, n, k, prec):
    a = torch.randn(m, k, device=dev)
    w = torch.randn(n, k, device=dev) / k ** 0.5
    b = torch.randn(n, device=dev)
    g = torch.randn(n, device=dev)
    be = torch.randn(n, device=dev)
    out = torch.full((m, n), float("nan"), device=dev)
    rc = lib.osteo_linear_gn_silu_tc(a.data_ptr(), w.data_ptr(), b.data_ptr(), g.data_ptr(), be.data_ptr(), out.data_ptr(), m, n, k, prec, None)
    if rc != 0:
        print(f"gn m={m} n={n} k={k} prec={prec}: ERROR {lib.osteo_last_error().decode()}", flush=True)
        return
    y = torch.nn.functional.linear(a.double(), w.double(), b.double())
    y = torch.nn.functional.group_norm(y, 8, g.double(), be.double(), 1e-5)
    ref = torch.nn.functional.silu(y).float()
    err = (out - ref).abs().max().item()
    rel = ((out - ref).norm() / ref.norm()).item()
    print(f"gn_silu m={m} n={n} k={k} prec={prec}: max_abs_err={err:.3e} rel_fro={rel:.3e} nan={torch.isnan(out).sum().item()}", flush=True)

for (m, n, k) in [(128, 128, 64), (333, 256, 256), (1000, 512, 512), (515, 512, 1024)]:
    for prec in (0, 1):
        run_gn(m, n, k, prec)

# throughput probe
for (m, n, k) in [(32768, 512, 512), (32768, 256, 5184)]:
    a = torch.randn(m, k, device=dev); w = torch.randn(n, k, device=dev); out = torch.empty(m, n, device=dev)
    lib.osteo_linear_tc(a.data_ptr(), w.data_ptr(), None, out.data_ptr(), m, n, k, 0, None)
    torch.cuda.synchronize(); t0 = time.time()
    lib.osteo_linear_tc(a.data_ptr(), w.data_ptr(), None, out.data_ptr(), m, n, k, 0, None)
    torch.cuda.synchronize(); print(f"linear_tc incl. pack m={m} n={n} k={k}: {1e3*(time.time()-t0):.2f} ms", flush=True)
print("DONE", flush=True)
