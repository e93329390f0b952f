"""GPU: BASELINE.json configs[0] through the drop-in -- the reference's UNMODIFIED Trainer (5 epochs, n = 100, batch 16, mixup),
load_trained_model on the checkpoint it wrote, SyntheticPatientGenerator.generate_scenarios and validate_all, all driven by the
reference's own main.py functions with `models.diffusion` / `utils.validation` resolving to the B200-native classes
(SURVEY.md §4 item 7, §8b). The reference sources are the staged copy oracle/_ref/reference (oracle/stage_reference.py)."""
import numpy as np
import pytest
import torch

from oracle import reference_import as R

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not R.available(), reason="reference sources not staged (python -m oracle.stage_reference)")]


def test_reference_pipeline_runs_unmodified_against_the_dropin(tmp_path):
    from tests.dropin_harness import run_pipeline

    history, synthetic, results, config, owners = run_pipeline(tmp_path, use_dropin=True, n=100, dims=(62, 5054, 26), epochs=5, n_generate=99)
    assert owners == ("osteosarcoma_diffusionmodel_b200.diffusion", "osteosarcoma_diffusionmodel_b200.validation")
    # utils/train.py:296-339: 5 epochs of 5 batches (80 train rows, batch 16, drop_last), validation after each
    assert len(history["train_loss"]) == 5 and len(history["val_loss"]) == 5
    assert all(np.isfinite(v) and 0.0 < v < 10.0 for v in history["train_loss"] + history["val_loss"])
    assert history["train_loss"][-1] < history["train_loss"][0] + 0.05          # eps-prediction MSE starts near 1 and does not blow up
    # utils/train.py:275-294 checkpoint layout, written through OUR state_dict, re-read by utils/generate.py:238-298 (strict load)
    ck = torch.load(tmp_path / "results" / "checkpoints" / "best_model.pt", map_location="cpu")
    assert set(ck) == {"epoch", "model_state_dict", "optimizer_state_dict", "val_loss", "config"}
    assert len(ck["model_state_dict"]) == 56 and len(ck["optimizer_state_dict"]["state"]) == 52
    # utils/generate.py:96-175 on three scenarios, 33 patients each
    assert set(synthetic) == {s["name"] for s in config["generation"]["scenarios"]}
    for d in synthetic.values():
        assert d["mutations"].shape == (33, 62) and d["expression"].shape == (33, 5054) and d["pathways"].shape == (33, 26) and d["conditions"].shape == (33, 3)
        assert set(np.unique(d["mutations"])) <= {0.0, 1.0}
        assert np.isfinite(d["expression"]).all() and np.isfinite(d["pathways"]).all()
    # utils/validation.py:300-387 via main.py:322
    for k in ("mutation_frequency_correlation", "cooccurrence_pattern_correlation", "mutual_exclusivity_violation_rate", "ks_test_mean_pvalue",
              "ks_test_fraction_significant", "mmd", "wasserstein_distance_mean", "overall_biological_score"):
        assert k in results, k
    assert np.isfinite(results["mmd"]) and 0.0 <= results["mmd"] <= 2.0 ** 0.5 + 1e-6
    assert (tmp_path / "results" / "validation_results.csv").exists()


def test_reference_written_checkpoint_round_trip(tmp_path):
    """A checkpoint written by the REFERENCE's Trainer around the REFERENCE's model (utils/train.py:275-294, one epoch on this GPU under
    PyTorch eager) loads through the reference's load_trained_model into the B200-native class (strict), its optimizer_state_dict loads
    into AdamW / FusedAdamW over our parameters, and one reverse step with injected noise matches the reference model holding the same
    weights to the fp32 tolerance."""
    import copy
    import os

    from tests.dropin_harness import pipeline_config, pipeline_modules, write_dummy_cohort
    from tests.helpers import TOL_BF16, TOL_FP32X3, rel

    dims = (30, 200, 10)
    write_dummy_cohort(tmp_path / "data" / "processed", 64, *dims, seed=3)
    (tmp_path / "config").mkdir()
    config = pipeline_config(R.REFERENCE_ROOT, tmp_path, 1, 6)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        with pipeline_modules(use_dropin=False) as main:          # the reference alone writes the checkpoint
            torch.manual_seed(1)
            main.train_model(copy.deepcopy(config))
            from utils.generate import load_trained_model
            ref_model = load_trained_model(tmp_path / "results" / "checkpoints" / "best_model.pt", copy.deepcopy(config), "cuda")
            assert type(ref_model).__module__ == "models.diffusion"
        with pipeline_modules(use_dropin=True):                   # ... and the drop-in reads it through the reference's loader
            from utils.generate import load_trained_model
            model = load_trained_model(tmp_path / "results" / "checkpoints" / "best_model.pt", copy.deepcopy(config), "cuda")
            assert type(model).__module__ == "osteosarcoma_diffusionmodel_b200.diffusion"
    finally:
        os.chdir(cwd)
    ck = torch.load(tmp_path / "results" / "checkpoints" / "best_model.pt", map_location="cuda")
    for k, v in model.state_dict().items():
        assert torch.equal(v, ck["model_state_dict"][k]), k
    from osteosarcoma_diffusionmodel_b200.optim import FusedAdamW
    for opt in (torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5), FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)):
        opt.load_state_dict(ck["optimizer_state_dict"])
        assert len(opt.state_dict()["state"]) == 52
    D = sum(dims)
    rs = np.random.RandomState(5)
    x = torch.from_numpy(rs.standard_normal((37, D)).astype(np.float32)).cuda()
    z = torch.from_numpy(rs.standard_normal((37, D)).astype(np.float32)).cuda()
    cond = torch.from_numpy(rs.standard_normal((37, 3)).astype(np.float32)).cuda()
    for t in (999, 500, 1):
        orig = torch.randn_like
        torch.randn_like = lambda *_a, **_k: z.clone()
        try:
            with torch.no_grad():
                ref = ref_model.p_sample(x, t, cond)
        finally:
            torch.randn_like = orig
        assert rel(model.set_precision("fp32x3").p_sample(x, t, cond, noise=z), ref) < TOL_FP32X3, t
        assert rel(model.set_precision("bf16").p_sample(x, t, cond, noise=z), ref) < TOL_BF16, t
    model.check_status()
