"""GPU: building-block kernels through the C-ABI against torch fp64 / the numpy Philox oracle."""
import numpy as np
import pytest
import torch

from oracle import philox_oracle as P
from osteosarcoma_diffusionmodel_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    return _lib.load()


def _linear(lib, m, n, k, prec, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) / k ** 0.5
    b = torch.randn(n, device="cuda", generator=g)
    out = torch.full((m, n), float("nan"), device="cuda")
    _lib.check(lib.osteo_linear_tc(a.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), m, n, k, prec, None))
    return a, w, b, out


@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (1, 128, 64), (100, 130, 70), (257, 512, 1024), (515, 256, 5142), (300, 5142, 256)])
def test_linear_tc_fp32x3(lib, m, n, k):
    a, w, b, out = _linear(lib, m, n, k, _lib.PREC_FP32X3)
    ref = (a.double() @ w.double().t() + b.double())
    assert not torch.isnan(out).any()
    assert ((out.double() - ref).norm() / ref.norm()).item() < 2e-5   # fp32 tolerance: rel 1e-4 (we are ~5x inside)


@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (100, 130, 70), (257, 512, 1024), (515, 256, 5142)])
def test_linear_tc_bf16_equals_bf16_rounded_inputs(lib, m, n, k):
    a, w, b, out = _linear(lib, m, n, k, _lib.PREC_BF16)
    ref = (a.bfloat16().double() @ w.bfloat16().double().t() + b.double())
    # exact products of bf16 inputs, fp32 accumulation: only summation-order noise remains
    assert (out.double() - ref).abs().max().item() < 1e-4 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (333, 256, 256), (1000, 512, 512), (515, 512, 1024)])
@pytest.mark.parametrize("prec,tol", [(_lib.PREC_FP32X3, 2e-5), (_lib.PREC_BF16, 5e-3)])
def test_linear_groupnorm_silu(lib, m, n, k, prec, tol):
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn(m, k, device="cuda", generator=g)
    w = torch.randn(n, k, device="cuda", generator=g) / k ** 0.5
    b, ga, be = (torch.randn(n, device="cuda", generator=g) for _ in range(3))
    out = torch.full((m, n), float("nan"), device="cuda")
    _lib.check(lib.osteo_linear_gn_silu_tc(a.data_ptr(), w.data_ptr(), b.data_ptr(), ga.data_ptr(), be.data_ptr(), out.data_ptr(), m, n, k, prec, None))
    y = torch.nn.functional.linear(a.double(), w.double(), b.double())
    ref = torch.nn.functional.silu(torch.nn.functional.group_norm(y, 8, ga.double(), be.double(), 1e-5))
    assert ((out.double() - ref).norm() / ref.norm()).item() < tol


def test_groupnorm_width_is_validated(lib):
    a = torch.zeros(8, 64, device="cuda")
    w = torch.zeros(192, 64, device="cuda")
    v = torch.zeros(192, device="cuda")
    o = torch.zeros(8, 192, device="cuda")
    assert lib.osteo_linear_gn_silu_tc(a.data_ptr(), w.data_ptr(), v.data_ptr(), v.data_ptr(), v.data_ptr(), o.data_ptr(), 8, 192, 64, 0, None) != 0


def test_philox_words_bit_exact(lib):
    n, ncol4, seed, row_base, stream, step = 37, 13, 0x1234567890ABCDEF, (1 << 33) + 5, 2, 999
    out = torch.zeros(n, ncol4 * 4, dtype=torch.int32, device="cuda")
    _lib.check(lib.osteo_philox_words(out.data_ptr(), n, ncol4, seed, row_base, stream, step, None))
    got = out.cpu().numpy().view(np.uint32).reshape(n, ncol4, 4)
    ref = P.words(seed, np.arange(n, dtype=np.uint64) + np.uint64(row_base), ncol4, stream, step)
    assert np.array_equal(got, ref)


def test_philox_normals_match_oracle_and_moments(lib):
    n, d, seed = 64, 5142, 42
    out = torch.zeros(n, d, device="cuda")
    _lib.check(lib.osteo_philox_normal(out.data_ptr(), n, d, seed, 7, 0, 123, None))
    ref = P.normals(seed, np.arange(n, dtype=np.uint64) + np.uint64(7), d, 0, 123)
    # Stated tolerance 2e-4 absolute on N(0,1) draws: the MUFU lg2 has ~2^-22 ABSOLUTE error, i.e. a relative error in
    # -ln(u1) that grows as u1 -> 1 (tiny radii); measured worst case 7e-5, typical 1e-7.
    err = np.abs(out.cpu().numpy().astype(np.float64) - ref)
    assert err.max() < 2e-4 and np.median(err) < 5e-7
    big = torch.zeros(4096, 2048, device="cuda")
    _lib.check(lib.osteo_philox_normal(big.data_ptr(), 4096, 2048, 1, 0, 1, 0, None))
    assert abs(big.mean().item()) < 2e-3 and abs(big.var().item() - 1.0) < 3e-3
    kurt = (big ** 4).mean().item()
    assert abs(kurt - 3.0) < 0.02
    # the PACKED mapping of the reverse-noise stream (stream 0: one word per Box-Muller pair, 20-bit radius / 12-bit angle)
    packed = torch.zeros(4096, 2048, device="cuda")
    _lib.check(lib.osteo_philox_normal(packed.data_ptr(), 4096, 2048, 1, 0, 0, 17, None))
    assert abs(packed.mean().item()) < 2e-3 and abs(packed.var().item() - 1.0) < 3e-3
    assert abs((packed ** 4).mean().item() - 3.0) < 0.02 and abs((packed ** 6).mean().item() - 15.0) < 0.3
    assert 4.5 < packed.abs().max().item() <= 5.2655        # sqrt(2 * 20 * ln 2)
    pairs = packed.view(4096, 1024, 2)
    assert abs((pairs[..., 0] * pairs[..., 1]).mean().item()) < 2e-3                      # the two normals of a pair are uncorrelated
    assert abs((packed[:, :-2] * packed[:, 2:]).mean().item()) < 2e-3                     # and so are neighbouring pairs
    ref8 = P.normals(1, np.arange(4, dtype=np.uint64), 2048, 0, 17)
    assert np.abs(packed[:4].cpu().numpy() - ref8).max() < 2e-4
    # distinct (stream, step, row_base) give distinct streams
    other = torch.zeros(4096, 2048, device="cuda")
    _lib.check(lib.osteo_philox_normal(other.data_ptr(), 4096, 2048, 1, 0, 1, 1, None))
    assert abs((big * other).mean().item()) < 2e-3
