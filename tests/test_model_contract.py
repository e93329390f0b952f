"""CPU: the drop-in class keeps the reference's constructor / attribute / state_dict contract."""
import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as O
from oracle import synth
from osteosarcoma_diffusionmodel_b200.diffusion import BiologyAwareDiffusionModel


def make(dims=synth.CONFIG_YAML_DIMS, **kw):
    return BiologyAwareDiffusionModel(dims["mutation_dim"], dims["expression_dim"], dims["pathway_dim"], dims["condition_dim"], synth.model_config(**kw))


def test_state_dict_keys_shapes_and_order():
    m = make()
    sd = m.state_dict()
    expect = synth.param_shapes(5142, 3, (256, 512, 256))
    names = [n for n, _ in expect] + ["betas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"]
    assert len(sd) == 56
    assert sorted(sd.keys()) == sorted(names)
    assert [n for n, _ in m.named_parameters()] == [n for n, _ in expect]      # reference registration order
    for n, shape in expect:
        assert tuple(sd[n].shape) == tuple(shape), n
    assert sum(p.numel() for p in m.parameters()) == 4_275_798
    assert sd["condition_embed.mlp.0.weight"].shape[1] == 3     # utils/generate.py:247 keys off this


def test_param_order_matches_c_abi_order():
    m = make(hidden_dims=(128, 256))
    names = {id(p): n for n, p in m.named_parameters()}
    assert [names[id(p)] for p in m._param_list()] == [n for n, _ in synth.param_shapes(5142, 3, (128, 256))]


def test_attributes_and_buffers(golden_dir):
    m = make()
    assert (m.mutation_dim, m.expression_dim, m.pathway_dim, m.condition_dim, m.data_dim, m.num_steps) == (62, 5054, 26, 3, 5142, 1000)
    assert not hasattr(m, "vae")            # utils/train.py:233 dispatches on this
    g = np.load(golden_dir / "ddpm_config.npz")
    for k in ("betas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"):
        assert np.array_equal(getattr(m, k).numpy(), g[f"buf_{k}"]), k
    lin = make(schedule="linear")
    assert torch.equal(lin.betas, torch.linspace(1e-4, 0.02, 1000))


def test_unknown_schedule_raises_like_reference():
    with pytest.raises(ValueError, match="Unknown schedule: sigmoid"):
        make(schedule="sigmoid")


def test_unknown_config_keys_are_tolerated():
    cfg = synth.model_config()
    cfg["model"]["constraints"] = {"pathway_coherence_weight": 1.0}
    cfg["generation"] = {"guidance_scale": 7.5}
    BiologyAwareDiffusionModel(10, 20, 5, 3, cfg)


def test_strict_load_of_reference_shaped_state_dict():
    m = make(dims=synth.SMOKE_DIMS)
    sd = synth.make_params(350, 5, (256, 512, 256), seed=1)
    sd.update(O.schedule_buffers("cosine", 1000))
    m.load_state_dict(sd, strict=True)
    assert torch.equal(m.unet.decoder[1][0].weight, sd["unet.decoder.1.0.weight"])


def test_time_embedding_table_matches_reference_rows(golden_dir):
    g = np.load(golden_dir / "ddpm_smoke.npz")
    tab = make(dims=synth.SMOKE_DIMS).unet.time_embed.table(1000)
    assert np.array_equal(tab[g["temb_rows"]].numpy(), g["temb"])


def test_reverse_coefficients_equal_oracle():
    m = make()
    a = m.reverse_coefficients(m.betas, m.alphas_cumprod)
    b = O.reverse_coefficients(m.betas, m.alphas_cumprod)
    for u, v in zip(a, b):
        assert np.array_equal(u, v)


def test_optimizer_and_clip_see_the_parameters():
    m = make(dims=synth.SMOKE_DIMS)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=1e-5)   # utils/train.py:169-173
    assert sum(len(g["params"]) for g in opt.param_groups) == 52


@pytest.mark.parametrize("hidden", [(256, 512, 256), (128, 256), (512, 512, 512, 256)])
def test_data_parallel_cut_is_a_half_block_boundary_near_the_middle(hidden):
    """model._dp_cut: the backward pass is cut before half block `cut`; the gradients finished by then are the tensors [10 + 4 cut, end)
    of the C-ABI order = elements [offset, total) of the flat gradient buffer -- a contiguous tail, about half of the bytes, that always
    contains output_proj (the first gradients the backward pass produces)."""
    m = make(hidden_dims=hidden)
    ps = m._param_list()
    sizes = [p.numel() for p in ps]
    cut, off = m._dp_cut(ps)
    n_halves = (len(ps) - 12) // 4
    assert len(ps) == 12 + 4 * n_halves and 1 <= cut <= n_halves
    assert off == sum(sizes[:10 + 4 * cut])
    total = sum(sizes)
    assert off + sizes[-1] + sizes[-2] <= total                 # output_proj (weight, bias) is in the tail
    assert 0.2 < off / total < 0.8
