"""Harness of the drop-in tests: BASELINE.json configs[0] -- the reference's OWN pipeline functions (main.py:112-340: train_model,
generate_synthetic_patients, validate_synthetic_patients, i.e. utils/train.py Trainer, utils/generate.py load_trained_model /
SyntheticPatientGenerator, utils/validation.py validate_all) run UNMODIFIED on a QUICKSTART-style dummy cohort (QUICKSTART.md:206-248)
of the preprocessor's output shape, with `models.diffusion` / `utils.validation` resolving either to the reference's modules
(use_dropin=False: the harness checks itself on the CPU) or to this repo's B200-native shims (use_dropin=True)."""
from __future__ import annotations

import copy
import importlib.util
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
NAMED_GENES = ["TP53", "RB1", "ATRX", "MDM2", "MYC"]


def write_dummy_cohort(processed: Path, n: int, n_mut: int, n_expr: int, n_path: int, seed: int = 0) -> None:
    import pandas as pd

    rs = np.random.RandomState(seed)
    processed.mkdir(parents=True, exist_ok=True)
    ids = [f"P{i}" for i in range(n)]
    mgenes = NAMED_GENES + [f"M{i}" for i in range(n_mut - len(NAMED_GENES))]
    pd.DataFrame(rs.randint(0, 2, (n, n_mut)), index=ids, columns=mgenes).to_csv(processed / "mutation_matrix_aligned.csv")
    pd.DataFrame(rs.standard_normal((n, n_expr)).astype(np.float32), index=ids, columns=[f"E{i}" for i in range(n_expr)]).to_csv(processed / "expression_matrix_aligned.csv")
    pnames = ["HALLMARK_P53_PATHWAY", "HALLMARK_MYC_TARGETS_V1"] + [f"HALLMARK_X{i}" for i in range(n_path - 2)]
    pd.DataFrame(rs.standard_normal((n, n_path)), index=ids, columns=pnames).to_csv(processed / "pathway_scores.csv")
    pd.DataFrame({"submitter_id": ids, "survival_days": rs.randint(100, 2000, n), "event_occurred": rs.randint(0, 2, n),
                  "age_years": rs.uniform(10, 18, n)}).to_csv(processed / "clinical_aligned.csv", index=False)


def pipeline_config(reference_root: str, work: Path, epochs: int, n_generate: int) -> dict:
    import yaml

    with open(os.path.join(reference_root, "config", "config.yaml")) as f:
        config = yaml.safe_load(f)
    config["data"]["processed_dir"] = str(work / "data" / "processed")
    config["training"]["num_epochs"] = epochs
    config["training"]["save_dir"] = str(work / "results" / "checkpoints")
    config["generation"]["num_synthetic_samples"] = n_generate
    config.setdefault("output", {})
    config["output"]["synthetic_data_dir"] = str(work / "results" / "synthetic")
    config["output"]["results_dir"] = str(work / "results")
    return config


class pipeline_modules:
    """Context: sys.path / sys.modules arranged so that the reference's main.py imports resolve either to the reference alone or to the
    repo's shims first. Yields the reference's `main` module."""

    def __init__(self, use_dropin: bool):
        self.use_dropin = use_dropin

    def __enter__(self):
        from oracle import reference_import as R

        self.R = R
        if not R.available():
            raise RuntimeError("reference not available")
        R._install_stub()
        self._path = list(sys.path)
        self._mods = {k: v for k, v in sys.modules.items() if k in ("models", "utils", "main", "data") or k.startswith(("models.", "utils.", "data."))}
        for k in self._mods:
            del sys.modules[k]
        spec = importlib.util.spec_from_file_location("main", os.path.join(R.REFERENCE_ROOT, "main.py"))
        main = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(main)           # main.py:14 puts the reference root first on sys.path ...
        import torch

        main.torch = torch                      # main.py:220 reads `torch.cuda.is_available()` but main.py never imports torch
        rest = [p for p in sys.path if p not in (str(ROOT), R.REFERENCE_ROOT, "")]
        # ... a drop-in user puts this repo in front of it
        sys.path[:] = ([str(ROOT), R.REFERENCE_ROOT] if self.use_dropin else [R.REFERENCE_ROOT, str(ROOT)]) + rest
        importlib.invalidate_caches()
        return main

    def __exit__(self, *exc):
        for k in [k for k in sys.modules if k in ("models", "utils", "main", "data") or k.startswith(("models.", "utils.", "data."))]:
            del sys.modules[k]
        sys.modules.update(self._mods)
        sys.path[:] = self._path


def run_pipeline(work: Path, use_dropin: bool, n: int = 100, dims=(62, 5054, 26), epochs: int = 5, n_generate: int = 99, seed: int = 0):
    """Returns (history, synthetic dict, validation results, config, model class module name)."""
    import torch
    from oracle import reference_import as R

    write_dummy_cohort(work / "data" / "processed", n, *dims, seed=seed)
    (work / "config").mkdir(parents=True, exist_ok=True)
    config = pipeline_config(R.REFERENCE_ROOT, work, epochs, n_generate)
    cwd = os.getcwd()
    os.chdir(work)
    try:
        with pipeline_modules(use_dropin) as main:
            torch.manual_seed(seed)
            np.random.seed(seed)
            history = main.train_model(copy.deepcopy(config))
            synthetic = main.generate_synthetic_patients(copy.deepcopy(config))
            results = main.validate_synthetic_patients(copy.deepcopy(config))
            from models.diffusion import BiologyAwareDiffusionModel
            from utils.validation import BiologicalValidator

            owners = (BiologyAwareDiffusionModel.__module__, BiologicalValidator.__module__)
    finally:
        os.chdir(cwd)
    return history, synthetic, results, config, owners
