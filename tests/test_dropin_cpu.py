"""CPU: the drop-in import layout (namespace packages merge with the reference's `models/` and `utils/`) and the pipeline harness
itself, run once against the REFERENCE's own classes at reduced dims (so tests/test_dropin_gpu.py exercises a harness known to work)."""
import pytest

from oracle import reference_import as R

needs_reference = pytest.mark.skipif(not R.available(), reason="reference sources not staged (python -m oracle.stage_reference)")


@needs_reference
def test_shims_merge_with_the_reference_namespace_packages():
    from tests.dropin_harness import ROOT

    with R.dropin_path(str(ROOT)):
        import models.cvae  # noqa: F401  (reference)
        import utils.generate  # noqa: F401
        import utils.pathway_features  # noqa: F401
        import utils.train  # noqa: F401
        from models.diffusion import BiologyAwareDiffusionModel
        from utils.validation import BiologicalValidator

        assert BiologyAwareDiffusionModel.__module__ == "osteosarcoma_diffusionmodel_b200.diffusion"
        assert BiologicalValidator.__module__ == "osteosarcoma_diffusionmodel_b200.validation"
        assert utils.train.__file__.startswith(R.REFERENCE_ROOT)
        for name in ("validate_all", "validate_mutation_cooccurrence", "statistical_tests", "compute_mmd", "validate_pathway_coherence",
                     "validate_mutation_expression_correlation"):
            assert callable(getattr(BiologicalValidator, name)), name


def test_no_init_files_shadow_the_reference_packages():
    from tests.dropin_harness import ROOT

    assert not (ROOT / "models" / "__init__.py").exists() and not (ROOT / "utils" / "__init__.py").exists()


@needs_reference
def test_harness_runs_the_reference_pipeline(tmp_path):
    from tests.dropin_harness import run_pipeline

    history, synthetic, results, config, owners = run_pipeline(tmp_path, use_dropin=False, n=40, dims=(12, 60, 8), epochs=2, n_generate=6)
    assert owners == ("models.diffusion", "utils.validation")
    assert len(history["train_loss"]) == 2 and len(history["val_loss"]) == 2
    assert set(synthetic) == {s["name"] for s in config["generation"]["scenarios"]}
    for d in synthetic.values():
        assert d["mutations"].shape == (2, 12) and d["expression"].shape == (2, 60) and d["pathways"].shape == (2, 8)
    assert "overall_biological_score" in results and "mmd" in results
    assert (tmp_path / "results" / "checkpoints" / "best_model.pt").exists()
