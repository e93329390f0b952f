"""CPU: host logic of the multi-task wrapper and the fused optimiser (no compute: both fail loudly without CUDA)."""
import pytest
import torch

from oracle import synth
from osteosarcoma_diffusionmodel_b200.multitask import BiologyConstrainedDiffusion, correlation_losses, pack_column_sets
from osteosarcoma_diffusionmodel_b200.optim import FusedAdamW


def _cfg():
    cfg = synth.model_config()
    cfg["model"]["constraints"] = {"pathway_coherence_weight": 2.0, "mutation_expression_weight": 0.25, "survival_prediction_weight": 0.1}
    return cfg


def test_wrapper_layout_follows_the_reference_wrapper():
    """models/cvae.py:222-346: same constructor arguments, survival head shapes and loss weights; the wrapped model is not called `vae`
    (utils/train.py:233 dispatches on that name); gathered columns = union of pathway genes, rule columns and the head's inputs."""
    m = BiologyConstrainedDiffusion(62, 5054, 26, 3, _cfg(), pathway_members=[[1, 2, 3, 4], [3, 10, 11]], correlation_rules=[(0, 1, -1), (5, 3, 1)])
    assert not hasattr(m, "vae") and hasattr(m, "diffusion")
    assert (m.pathway_coherence_weight, m.mutation_expr_weight, m.survival_weight) == (2.0, 0.25, 0.1)
    sd = m.state_dict()
    assert sd["survival_predictor.0.weight"].shape == (128, 128) and sd["survival_predictor.3.weight"].shape == (1, 128)
    assert len([k for k in sd if k.startswith("diffusion.")]) == 56          # 52 parameters + 4 schedule buffers
    cols = m._aux_columns.tolist()
    assert cols == sorted(set(cols))                                          # train_inject requires distinct columns
    expr0, path0 = 62, 62 + 5054
    for g in (1, 2, 3, 4, 10, 11):
        assert expr0 + g in cols
    assert 0 in cols and 5 in cols and path0 + 1 in cols and path0 + 3 in cols
    # sets index INTO the gathered matrix
    assert [cols[i] for i in m._sets[0]] == [expr0 + 1, expr0 + 2, expr0 + 3, expr0 + 4]
    assert [cols[i] for i in m._sets[2]] == [0, path0 + 1] and m._modes == [0, 0, -1, 1]
    # the head reads the mutation block then the pathway-score block, truncated to latent_dim
    head_cols = [cols[i] for i in m._head_pos.tolist()]
    assert head_cols == list(range(62)) + list(range(path0, path0 + 26))


@pytest.mark.parametrize("kw", [dict(pathway_members=[[1, 2]]), dict(pathway_members=[list(range(40))]), dict(pathway_members=[[1, 2, 6000]]),
                                dict(correlation_rules=[(62, 0, 1)]), dict(correlation_rules=[(0, 26, 1)]), dict(correlation_rules=[(0, 0, 2)])])
def test_wrapper_rejects_bad_sets(kw):
    with pytest.raises(ValueError):
        BiologyConstrainedDiffusion(62, 5054, 26, 3, _cfg(), **kw)


def test_no_cpu_fallback():
    m = BiologyConstrainedDiffusion(20, 90, 10, 2, synth.model_config(hidden_dims=(128, 256)), pathway_members=[[0, 1, 2]]).train()
    with pytest.raises(RuntimeError):
        m(torch.zeros(4, 120), torch.zeros(4, 2))
    with pytest.raises(RuntimeError):
        correlation_losses(torch.zeros(8, 4), [[0, 1]], [0])
    p = torch.nn.Parameter(torch.zeros(3))
    p.grad = torch.ones(3)
    with pytest.raises(RuntimeError):
        FusedAdamW([p]).step()
    with pytest.raises(ValueError):
        FusedAdamW([p], lr=-1.0)


def test_pack_column_sets_groups_of_32_and_checks():
    packs = pack_column_sets([[i, i + 1, i + 2] for i in range(40)], [0] * 40, "cpu")
    assert [p[0].shape[0] for p in packs] == [32, 8]
    assert packs[0][0].dtype == torch.int32 and packs[0][0][0].tolist()[:4] == [0, 1, 2, -1]
    with pytest.raises(ValueError):
        pack_column_sets([[0]], [0], "cpu")
    with pytest.raises(ValueError):
        pack_column_sets([[0, 1, 2]], [1], "cpu")
    with pytest.raises(ValueError):
        pack_column_sets([[0, 1]], [0, 1], "cpu")


def test_ingress_has_no_cpu_fallback_and_checks_shapes():
    from osteosarcoma_diffusionmodel_b200.ingress import GpuResidentDataset, MixupAugmentation
    x, c, sv = torch.zeros(8, 6), torch.zeros(8, 2), torch.zeros(8)
    with pytest.raises(ValueError):
        GpuResidentDataset(x, c[:4], sv, device="cpu")
    ds = GpuResidentDataset(x, c, sv, device="cpu")            # construction is plain tensor plumbing ...
    assert len(ds) == 8
    with pytest.raises(RuntimeError):                          # ... every batch is a kernel of the library
        ds.gather(torch.arange(4))
    with pytest.raises(RuntimeError):
        MixupAugmentation(0.2)({"data": x, "conditions": c, "survival": sv}, lam=0.5, index=torch.arange(8))
