"""ORACLE support — the 'production-mode' parity case shared by oracle/gen_golden.py and the GPU tests: a cohort big enough for the
benchmarked configuration (bf16, fused step kernel, replayed graphs, two parallel row branches, in-kernel Philox) of which a few
hundred rows -- first tile, the tile pair around the branch boundary, the ragged last tile -- are also pushed through the REFERENCE's
own `sample()` with the same x_T / z (the Philox streams restated in oracle/philox_oracle.py). Test infrastructure."""
from __future__ import annotations

import numpy as np

N_TOTAL = 76_100          # 595 row tiles of 128 >= 2 branches x 2 waves x 148 SMs (csrc/osteo_ddpm.cu: capture_steps); last tile holds 68 rows
SEED = 20261018
PARAM_SEED = 2            # the weights of the 'config' fixture (oracle/synth.py: make_params)
STREAM_REVERSE, STREAM_XT = 0, 1          # csrc/philox.cuh


def rows() -> np.ndarray:
    """208 global row indices: [0, 64) | [37 980, 38 060) across the branch boundary (tile 297 = row 38 016) | the last 64 rows."""
    return np.concatenate([np.arange(0, 64), np.arange(37_980, 38_060), np.arange(N_TOTAL - 64, N_TOTAL)]).astype(np.int64)


def stored_columns(mutation_dim: int = 62, expression_dim: int = 5054, pathway_dim: int = 26) -> np.ndarray:
    """Columns of the final sample kept in the fixture: every mutation and pathway column, every 8th expression column."""
    d = mutation_dim + expression_dim + pathway_dim
    return np.concatenate([np.arange(mutation_dim), np.arange(mutation_dim, mutation_dim + expression_dim, 8), np.arange(mutation_dim + expression_dim, d)])
