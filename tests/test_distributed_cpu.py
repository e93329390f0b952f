"""CPU, world_size 2 over gloo: the host-side sharding / reduction logic of the multi-GPU paths.
The CUDA partial-sum calls are replaced by the numpy oracle so the test checks exactly what the ranks exchange."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from osteosarcoma_diffusionmodel_b200 import distributed as D


def test_shard_rows_tiles_the_range_exactly():
    for n in (0, 1, 127, 128, 129, 1000, 100_000, 10_000_000):
        for ws in (1, 2, 3, 4, 8):
            for align in (1, 128):
                prev = 0
                for r in range(ws):
                    b, e = D.shard_rows(n, r, ws, align)
                    assert b == prev and b <= e <= n
                    if b < e:
                        assert b % align == 0
                    prev = e
                assert prev == n
    with pytest.raises(ValueError):
        D.shard_rows(10, 2, 2)


def test_buckets_are_reverse_ordered_and_capped():
    ts = [torch.zeros(s) for s in (10, 1000, 20, 300, 5)]
    b = D.make_buckets(ts, bucket_bytes=4 * 400)
    assert [i for bk in b for i in bk] == [4, 3, 2, 1, 0]
    assert b == [[4, 3, 2], [1], [0]]
    assert D.make_buckets(ts, bucket_bytes=1 << 30) == [[4, 3, 2, 1, 0]]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _init(rank, ws, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)


def _worker_grads(rank, ws, port, q):
    _init(rank, ws, port)
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    for i, p in enumerate(lin.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    nb = D.allreduce_gradients(lin.parameters(), bucket_bytes=64)
    ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(lin.parameters()))
    # dp_train_step keeps replicas identical: same averaged grads -> same clipped step
    opt = torch.optim.AdamW(lin.parameters(), lr=1e-2)

    class Wrap(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x, c, return_loss=True):
            return (self.m(x) - c).pow(2).mean()

    w = Wrap(lin)
    g = torch.Generator().manual_seed(100 + rank)
    x, c = torch.randn(16, 7, generator=g), torch.randn(16, 3, generator=g)
    D.dp_train_step(w, opt, x, c)
    flat = torch.cat([p.detach().reshape(-1) for p in lin.parameters()])
    gathered = [torch.zeros_like(flat) for _ in range(ws)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], t) for t in gathered)
    if rank == 0:
        q.put((ok, nb, same))
    dist.destroy_process_group()


def test_bucketed_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_grads, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    ok, nb, same = q.get(timeout=10)
    assert ok and nb >= 2 and same


def _worker_mmd(rank, ws, port, q):
    _init(rank, ws, port)
    from oracle import validators_oracle as V
    from osteosarcoma_diffusionmodel_b200 import validation as val

    rs = np.random.RandomState(5)
    n, m, d = 300, 260, 24
    X = rs.standard_normal((n, d)).astype(np.float32)
    Y = (rs.standard_normal((m, d)) + 0.3).astype(np.float32)

    def fake_partial(Xt, Yt, gamma, center, rx, ry, precision):
        Xn, Yn = Xt.numpy().astype(np.float64), Yt.numpy().astype(np.float64)
        sxx = np.exp(-gamma * V.sqeuclidean(Xn[rx[0]:rx[1]], Xn)).sum() if rx[1] > rx[0] else 0.0
        syy = np.exp(-gamma * V.sqeuclidean(Yn[ry[0]:ry[1]], Yn)).sum() if ry[1] > ry[0] else 0.0
        sxy = np.exp(-gamma * V.sqeuclidean(Xn[rx[0]:rx[1]], Yn)).sum() if rx[1] > rx[0] else 0.0
        return torch.tensor([sxx, syy, sxy], dtype=torch.float64)

    def fake_partial_cyclic(Xt, Yt, gamma, center, rnk, world, precision):
        """Host restatement of osteo_mmd_partial_cyclic: 128-row blocks b % world == rnk; Kxx / Kyy as symmetric half-Grams (128-column
        halves below the diagonal skipped, above it counted twice)."""
        Xn, Yn = Xt.numpy().astype(np.float64), Yt.numpy().astype(np.float64)

        def half_gram(A):
            nb, tot = -(-A.shape[0] // 128), 0.0
            for b in range(rnk, nb, world):
                K = np.exp(-gamma * V.sqeuclidean(A[128 * b:128 * b + 128], A))
                for hb in range(nb):
                    w = 2.0 if hb > b else (1.0 if hb == b else 0.0)
                    tot += w * K[:, 128 * hb:128 * hb + 128].sum()
            return tot

        sxy = sum(np.exp(-gamma * V.sqeuclidean(Xn[128 * b:128 * b + 128], Yn)).sum() for b in range(rnk, -(-Xn.shape[0] // 128), world))
        return torch.tensor([half_gram(Xn), half_gram(Yn), sxy], dtype=torch.float64)

    def fake_moments(data, cols, shift, rows):
        a = data.numpy().astype(np.float64)[rows[0]:rows[1]][:, list(cols)] - shift.numpy().astype(np.float64)
        k = len(cols)
        return torch.from_numpy(np.concatenate([[a.shape[0]], a.sum(0), (a.T @ a).reshape(-1)]))

    def fake_moments_batched(data, ci_t, shift, rows):
        a = data.numpy().astype(np.float64)[rows[0]:rows[1]]
        out = np.zeros((ci_t.shape[0], val._CM_STRIDE))
        for p, (ci, sh) in enumerate(zip(ci_t.numpy(), shift.numpy().astype(np.float64))):
            k = int((ci >= 0).sum())
            g = a[:, ci[:k]] - sh[:k]
            out[p, 0] = g.shape[0]
            out[p, 1:1 + k] = g.sum(0)
            s2 = np.zeros((32, 32))
            s2[:k, :k] = g.T @ g
            out[p, 33:] = s2.reshape(-1)
        return torch.from_numpy(out)

    def fake_moments_tiled(data, ci_t, rows, max_set_size=32, out=None):
        res = fake_moments_batched(data, ci_t, data[0, ci_t.clamp(min=0).long()], rows)
        if out is None:
            return res
        out.copy_(res)
        return out

    def fake_finish(mom, ci_t):
        out = []
        for blk, ci in zip(mom.numpy(), ci_t.numpy()):
            k = int((ci >= 0).sum())
            m = np.concatenate([blk[:1], blk[1:1 + k], blk[33:].reshape(32, 32)[:k, :k].reshape(-1)])
            out.append(val._corr_from_moments(m, k)[np.triu_indices(k, k=1)].mean())
        return torch.tensor(out, dtype=torch.float64)

    val._gram_partial_sums = fake_partial
    val._gram_partial_sums_cyclic = fake_partial_cyclic
    val._moments = fake_moments
    val._moments_batched = fake_moments_batched
    val._moments_tiled = fake_moments_tiled
    val._coherence_finish = fake_finish
    v = val.BiologicalValidator({"evaluation": {}}, device="cpu")
    v._require_cuda = lambda: None
    got = v.compute_mmd(X, Y)
    ref = V.compute_mmd(X, Y)
    members = [[0, 1, 2, 3], [4, 5, 6], [7, 8]]
    coh = v.pathway_coherence_from_tensors(torch.from_numpy(X), torch.from_numpy(Y[:, :d]), members)
    ref_coh = V.pathway_coherence(X, Y, members)
    if rank == 0:
        q.put((got, ref, coh, ref_coh))
    dist.destroy_process_group()


def test_mmd_and_coherence_row_sharding_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_mmd, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    got, ref, coh, ref_coh = q.get(timeout=10)
    assert abs(got - ref) < 1e-12
    for k in ref_coh:
        assert abs(coh[k] - ref_coh[k]) < 1e-10


class _FakeModel:
    data_dim = 4

    def __init__(self):
        self.calls = []

    def sample(self, cond, n, seed=0, row_base=0):
        self.calls.append((n, seed, row_base))
        base = torch.arange(row_base, row_base + n, dtype=torch.float32)[:, None]
        return base + cond[:, :1] * 0 + torch.zeros(n, self.data_dim)


def _worker_sample(rank, ws, port, q):
    _init(rank, ws, port)
    m = _FakeModel()
    cond = torch.zeros(11, 3)
    local = D.sample_sharded(m, cond, 11, seed=9)
    full = D.sample_sharded(m, cond, 11, seed=9, gather=True)
    if rank == 0:
        q.put((m.calls, local[:, 0].tolist(), full[:, 0].tolist()))
    dist.destroy_process_group()


def test_sample_sharded_world2_covers_every_global_row_once():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sample, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    calls, local, full = q.get(timeout=10)
    assert calls[0] == (6, 9, 0)
    assert local == [0, 1, 2, 3, 4, 5]
    assert full == list(range(11))


def _worker_shards(rank, ws, port, q, out_dir):
    _init(rank, ws, port)
    from tests.test_egress_cpu import _FakeModel
    n = 37
    cond = torch.arange(n * 2, dtype=torch.float32).reshape(n, 2)
    man = D.sample_sharded_to_shards(_FakeModel(), cond, n, out_dir, seed=2, rows_per_shard=8)
    if rank == 0:
        q.put(man["rows"])
    dist.destroy_process_group()


def test_sample_sharded_to_shards_world2_writes_the_whole_cohort_once(tmp_path):
    """Two ranks stream their row ranges to rank_000 / rank_001: together exactly the single-process cohort."""
    from osteosarcoma_diffusionmodel_b200.egress import load_shards
    from tests.test_egress_cpu import _FakeModel
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_shards, args=(r, 2, port, q, str(tmp_path))) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    rows0 = q.get(timeout=10)
    a, b = load_shards(tmp_path / "rank_000"), load_shards(tmp_path / "rank_001")
    assert rows0 == a["expression"].shape[0] and a["expression"].shape[0] + b["expression"].shape[0] == 37
    n = 37
    cond = torch.arange(n * 2, dtype=torch.float32).reshape(n, 2)
    ref = _FakeModel().sample_components(cond, n, seed=2, row_base=0)
    for k in ("expression", "pathways", "conditions"):
        assert np.array_equal(np.concatenate([a[k], b[k]]), ref[k].numpy())
    assert np.array_equal(np.concatenate([a["mutations"], b["mutations"]]), ref["mutations"].numpy().astype(float))
