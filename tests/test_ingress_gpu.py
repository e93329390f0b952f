"""GPU: device-resident dataset + fused gather/mixup against the torch-CPU restatement of utils/train.py:77-126,
bit-exact, including odd row pitches, 1-D survival, duplicate indices and the epoch iterator."""
import numpy as np
import pytest
import torch

from oracle import ingress_oracle as I
from oracle import synth
from osteosarcoma_diffusionmodel_b200.ingress import GpuResidentDataset, MixupAugmentation

pytestmark = pytest.mark.gpu


def _dataset(n, dims, seed=0):
    x0, cond = synth.make_cohort(n, *dims, seed=seed)
    surv = torch.from_numpy(np.random.RandomState(seed).gamma(2.0, 400.0, n).astype(np.float32))
    return x0, cond, surv


@pytest.mark.parametrize("n,dims,lam", [(300, (62, 5054, 26, 3), 0.37), (64, (20, 90, 11, 2), 0.9123456789), (33, (5, 7, 1, 1), 1.0), (17, (4, 4, 4, 3), 0.0)])
def test_gather_and_mixup_are_bit_exact(n, dims, lam):
    x0, cond, surv = _dataset(n, dims)
    ds = GpuResidentDataset(x0, cond, surv)
    rs = np.random.RandomState(1)
    b = max(1, n // 3)
    index = torch.from_numpy(rs.randint(0, n, size=b).astype(np.int64))           # duplicates allowed
    perm = torch.from_numpy(rs.permutation(b).astype(np.int64))
    ref_batch = I.collate(x0, cond, surv, index)
    got = ds.gather(index.cuda())
    for k in ref_batch:
        assert torch.equal(got[k].cpu(), ref_batch[k]), k
    ref_mix = I.mixup(ref_batch, lam, perm)
    mix = MixupAugmentation(alpha=0.2)
    fused = mix.gather(ds, index.cuda(), lam=lam, perm=perm.cuda())
    staged = mix({k: v.cuda() for k, v in ref_batch.items()}, lam=lam, index=perm.cuda())
    for k in ref_mix:
        assert torch.equal(fused[k].cpu(), ref_mix[k]), k
        assert torch.equal(staged[k].cpu(), ref_mix[k]), k
        assert fused[k].shape == ref_mix[k].shape


def test_epoch_iterator_matches_a_dataloader_epoch():
    """shuffle=False, drop_last=False: the batches of DataLoader(dataset, batch_size) in order; shuffle + drop_last: every row at most
    once, n // batch_size full batches; with mixup the batches feed the model's training step."""
    n, dims = 103, (20, 90, 10, 2)
    x0, cond, surv = _dataset(n, dims, seed=3)
    ds = GpuResidentDataset(x0, cond, surv)
    seen = 0
    for i, b in enumerate(ds.batches(16, shuffle=False, drop_last=False)):
        lo, hi = 16 * i, min(16 * i + 16, n)
        assert torch.equal(b["data"].cpu(), x0[lo:hi]) and torch.equal(b["survival"].cpu(), surv[lo:hi])
        seen += hi - lo
    assert seen == n
    g = torch.Generator(device="cuda").manual_seed(4)
    batches = list(ds.batches(16, shuffle=True, drop_last=True, generator=g))
    assert len(batches) == n // 16 and all(b["data"].shape == (16, 120) for b in batches)
    rows = torch.cat([b["survival"] for b in batches]).cpu()
    assert len(set(rows.tolist())) == rows.numel()            # gamma draws are distinct: no row twice
    from tests.helpers import build_model, load_case
    model = build_model(load_case("linear3"), "bf16").train()
    np.random.seed(0)
    for b in ds.batches(32, mixup=MixupAugmentation(0.2)):
        loss = model(b["data"], b["conditions"])
        loss.backward()
        assert torch.isfinite(loss)
    model.check_status()


def test_argument_checks():
    x0, cond, surv = _dataset(10, (4, 4, 4, 3))
    with pytest.raises(ValueError):
        GpuResidentDataset(x0, cond[:5], surv)
    ds = GpuResidentDataset(x0, cond, surv)
    with pytest.raises(ValueError):
        ds.gather(torch.arange(4, dtype=torch.int32, device="cuda"))
    with pytest.raises(ValueError):
        ds.gather(torch.arange(4))                # index on the host


def test_from_frames_on_the_device_feeds_the_batch_kernel():
    """The table preparation (prepare_data's normalisation + the dataset's alignment) done on the GPU equals the same on the CPU to fp32
    rounding, and its tensors go straight into the gather + mixup kernel."""
    import numpy as np
    from tests.test_ingress_frames_cpu import _tables
    from osteosarcoma_diffusionmodel_b200.ingress import GpuResidentDataset, MixupAugmentation

    mut, expr, path, clin = _tables(n=300, seed=8)
    gpu = GpuResidentDataset.from_frames(mut, expr, path, clin, device="cuda")
    cpu = GpuResidentDataset.from_frames(mut, expr, path, clin, device="cpu")
    assert gpu.data.is_cuda and len(gpu) == len(cpu) == 299
    np.testing.assert_allclose(gpu.data.cpu().numpy(), cpu.data.numpy(), rtol=2e-7, atol=1e-7)
    np.testing.assert_allclose(gpu.conditions.cpu().numpy(), cpu.conditions.numpy(), rtol=2e-7, atol=1e-7)
    assert torch.equal(gpu.survival_days.cpu(), cpu.survival_days)
    seen = 0
    for batch in gpu.batches(64, shuffle=True, drop_last=True, mixup=MixupAugmentation(0.2), generator=torch.Generator(device="cuda").manual_seed(1)):
        assert batch["data"].shape == (64, 7 + 11 + 5) and batch["conditions"].shape == (64, 3) and torch.isfinite(batch["data"]).all()
        seen += 1
    assert seen == 299 // 64
