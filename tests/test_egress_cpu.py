"""CPU: host logic of the shard writer / reader (egress.py) with a stand-in model; the kernels it drives are covered by the GPU suite."""
import json

import numpy as np
import torch

from osteosarcoma_diffusionmodel_b200.egress import generate_to_shards, load_shards


class _FakeModel:
    mutation_dim, expression_dim, pathway_dim = 11, 7, 3

    def sample_components(self, cond, n, seed=0, row_base=0, pack_bits=False):
        rows = torch.arange(row_base, row_base + n, dtype=torch.float32)[:, None]
        mut = ((rows.long() + torch.arange(self.mutation_dim)[None, :]) % 3 == 0).to(torch.uint8)
        out = {"mutations": mut, "expression": rows + torch.arange(self.expression_dim)[None, :] * 0.5 + seed,
               "pathways": -rows + torch.arange(self.pathway_dim)[None, :], "conditions": cond}
        if pack_bits:
            out["mutation_bits"] = torch.from_numpy(np.packbits(mut.numpy(), axis=1, bitorder="little"))
        return out


def test_shards_round_trip_and_do_not_depend_on_the_shard_size(tmp_path):
    model, n = _FakeModel(), 23
    cond = torch.arange(n * 2, dtype=torch.float32).reshape(n, 2)
    ref = model.sample_components(cond, n, seed=4, row_base=100)
    got = {}
    for shard_rows, bits in [(5, True), (23, False), (100, True)]:
        d = tmp_path / f"s{shard_rows}"
        man = generate_to_shards(model, cond, d, shard_rows=shard_rows, seed=4, row_base=100, pack_bits=bits)
        assert man["rows"] == n and sum(s["rows"] for s in man["shards"]) == n
        assert man["shards"][0]["row_begin"] == 100 and len(man["shards"]) == -(-n // shard_rows)
        assert json.loads((d / "manifest.json").read_text()) == man
        got[shard_rows] = load_shards(d)
    for g in got.values():
        assert g["mutations"].dtype == np.float64 and np.array_equal(g["mutations"], ref["mutations"].numpy().astype(float))
        assert np.array_equal(g["expression"], ref["expression"].numpy()) and np.array_equal(g["pathways"], ref["pathways"].numpy())
        assert np.array_equal(g["conditions"], cond.numpy())
    part = load_shards(tmp_path / "s5", shards=[1, 3])
    assert part["expression"].shape == (10, 7) and part["expression"][0, 0] == 105 + 4


def test_writer_failures_surface_and_the_manifest_lists_only_written_shards(tmp_path, monkeypatch):
    """A failure in the writer thread (full disk, permissions) must not be swallowed: generate_to_shards raises and never writes a
    manifest that advertises missing shards."""
    import pytest

    from osteosarcoma_diffusionmodel_b200 import egress

    model, n = _FakeModel(), 20
    cond = torch.zeros(n, 2)
    real_save = np.save
    calls = {"n": 0}

    def flaky_save(path, arr):
        calls["n"] += 1
        if "shard_00002" in str(path):
            raise OSError("No space left on device")
        return real_save(path, arr)

    monkeypatch.setattr(egress.np, "save", flaky_save)
    with pytest.raises(OSError, match="No space left"):
        generate_to_shards(model, cond, tmp_path / "x", shard_rows=5, pack_bits=False)
    assert not (tmp_path / "x" / "manifest.json").exists()
