"""GPU mirror of the three validators on the hot path: `BiologicalValidator.compute_mmd`,
`.validate_pathway_coherence`, `.validate_mutation_expression_correlation`
(utils/validation.py:273-298, :125-175, :177-223).  Same method names, argument meaning and return
values as the reference; the arithmetic runs in the C-ABI library (RBF Gram tiles on tcgen05 with an
exp/sum epilogue; column-gathered fp64 moment reduction for the Pearson correlations).  Under
torch.distributed the Gram rows / cohort rows are sharded over ranks and only the partial sums are
all-reduced.  No CPU fallback.
"""
from __future__ import annotations

import logging
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from . import distributed as D

logger = logging.getLogger(__name__)
_PRECISIONS = {"bf16": _lib.PREC_BF16, "fp32x3": _lib.PREC_FP32X3}


def _to_device(a, device) -> torch.Tensor:
    """DataFrame / ndarray / tensor -> contiguous fp32 tensor on `device`."""
    if hasattr(a, "values") and not isinstance(a, torch.Tensor):
        a = a.values
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(np.array(a, copy=True) if not a.flags.writeable else np.ascontiguousarray(a))
    return a.to(device=device, dtype=torch.float32).contiguous()


def _gram_partial_sums(X: torch.Tensor, Y: torch.Tensor, gamma: float, center: torch.Tensor, rx, ry, precision: int) -> torch.Tensor:
    """{sum Kxx, sum Kyy, sum Kxy} over Gram rows rx = [b, e) of X and ry of Y (fp64 tensor on X's device)."""
    sums = torch.zeros(3, dtype=torch.float64, device=X.device)
    _lib.check(_lib.load().osteo_mmd_partial(X.data_ptr(), X.shape[0], Y.data_ptr(), Y.shape[0], X.shape[1], float(gamma), center.data_ptr(),
                                             rx[0], rx[1], ry[0], ry[1], precision, sums.data_ptr(), _lib.stream_handle()))
    return sums


def _moments(data: torch.Tensor, cols: Sequence[int], shift: torch.Tensor, rows) -> torch.Tensor:
    """fp64 [1 + k + k*k] = {count, sum (x - s), sum (x - s)(x - s)^T} of the gathered columns over rows [b, e)."""
    k = len(cols)
    out = torch.zeros(1 + k + k * k, dtype=torch.float64, device=data.device)
    ci = torch.tensor(list(cols), dtype=torch.int32, device=data.device)
    _lib.check(_lib.load().osteo_corr_moments(data.data_ptr(), data.shape[0], data.stride(0), ci.data_ptr(), k, shift.data_ptr(), rows[0], rows[1],
                                              out.data_ptr(), _lib.stream_handle()))
    return out


_CM_STRIDE = 1 + 32 + 32 * 32      # osteo_corr_moments_batched: {count, s1[32], s2[32][32]} per column set


def _moments_batched(data: torch.Tensor, ci_t: torch.Tensor, shift: torch.Tensor, rows) -> torch.Tensor:
    """fp64 [n_sets, _CM_STRIDE]: the shifted moments of up to 32 gathered columns per set (ci_t int32 [n_sets, 32], -1 = unused
    slot) over rows [b, e), all sets in one pass over the cohort."""
    out = torch.empty((ci_t.shape[0], _CM_STRIDE), dtype=torch.float64, device=data.device)
    _lib.check(_lib.load().osteo_corr_moments_batched(data.data_ptr(), data.shape[0], data.stride(0), data.shape[1], ci_t.data_ptr(), ci_t.shape[0],
                                                      shift.data_ptr(), rows[0], rows[1], out.data_ptr(), _lib.stream_handle()))
    return out


def _corr_from_moments(mom: np.ndarray, k: int) -> np.ndarray:
    """Pearson correlation matrix from shifted moments (float64, host; k <= 32)."""
    n = mom[0]
    s1 = mom[1:1 + k]
    s2 = mom[1 + k:].reshape(k, k)
    cov = s2 - np.outer(s1, s1) / n
    sd = np.sqrt(np.diag(cov))
    with np.errstate(divide="ignore", invalid="ignore"):
        return cov / np.outer(sd, sd)


class BiologicalValidator:
    """Validate synthetic patients against biological knowledge (GPU-resident hot-path subset)."""

    def __init__(self, config: dict, device: Optional[str] = None, precision: str = "fp32x3"):
        self.config = config
        ev = config.get("evaluation", {})
        self.driver_genes = ev.get("driver_genes", [])
        self.mutually_exclusive_pairs = ev.get("mutually_exclusive_pairs", [])
        self.required_correlations = ev.get("required_correlations", [])
        if precision not in _PRECISIONS:
            raise ValueError(f"unknown precision {precision!r}")
        self.precision = precision
        self._index_cache: Dict[tuple, list] = {}          # (device, column sets) -> device index tensors of the moment kernel
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)

    def _require_cuda(self):
        if self.device.type != "cuda" or not torch.cuda.is_available():
            raise RuntimeError("BiologicalValidator (B200-native) computes only on a CUDA device; there is no CPU fallback")

    # ------------------------------------------------------------------ utils/validation.py:273-298
    def compute_mmd(self, X, Y, kernel: str = "rbf", gamma: Optional[float] = None) -> float:
        """sqrt(max(mean Kxx + mean Kyy - 2 mean Kxy, 0)) with K = exp(-gamma ||a - b||^2), diagonals included; gamma
        defaults to 1 / n_features; `kernel` is ignored exactly as in the reference."""
        self._require_cuda()
        X = _to_device(X, self.device)
        Y = _to_device(Y, self.device)
        if X.dim() != 2 or Y.dim() != 2 or X.shape[1] != Y.shape[1]:
            raise ValueError("X and Y must be 2-D with the same number of features")
        n, m, d = X.shape[0], Y.shape[0], X.shape[1]
        if gamma is None:
            gamma = 1.0 / d
        # RBF is translation invariant: centre on the pooled mean before the bf16 split (SURVEY.md §7 "MMD precision")
        center = ((X.sum(0, dtype=torch.float64) + Y.sum(0, dtype=torch.float64)) / (n + m)).float().contiguous()
        rank, ws = D.world()
        rx = D.shard_rows(n, rank, ws, align=128)
        ry = D.shard_rows(m, rank, ws, align=128)
        sums = _gram_partial_sums(X, Y, gamma, center, rx, ry, _PRECISIONS[self.precision])
        D.all_reduce_sum_(sums)
        sxx, syy, sxy = (float(v) for v in sums.cpu())
        mmd = sxx / (float(n) * n) + syy / (float(m) * m) - 2.0 * sxy / (float(n) * m)
        return float(np.sqrt(max(mmd, 0.0)))

    # ------------------------------------------------------------------ utils/validation.py:125-175
    def _coherence_moments(self, data: torch.Tensor, member_cols: List[List[int]]) -> torch.Tensor:
        """Device tensor [n_sets, _CM_STRIDE] of all-reduced moment blocks: all pathways in one pass over the cohort
        (osteo_corr_moments_batched: warp p owns pathway p, whole rows are streamed through shared memory). Nothing here waits for the
        device: the index tensors are cached per (device, column sets) and the result stays in HBM."""
        rank, ws = D.world()
        rows = D.shard_rows(data.shape[0], rank, ws)
        if any(len(c) > 32 for c in member_cols):
            raise ValueError("a pathway with more than 32 member genes is not supported by the moment kernel")
        key = (str(data.device), tuple(tuple(c) for c in member_cols))
        packs = self._index_cache.get(key)
        if packs is None:
            packs = []
            for b0 in range(0, len(member_cols), 32):
                sets = member_cols[b0:b0 + 32]
                ci = np.full((len(sets), 32), -1, dtype=np.int32)
                for i, cols in enumerate(sets):
                    ci[i, :len(cols)] = cols
                ci_t = torch.from_numpy(ci).to(data.device)
                packs.append((ci_t, ci_t.clamp(min=0).long()))
            if len(self._index_cache) > 16:
                self._index_cache.clear()
            self._index_cache[key] = packs
        parts = []
        for ci_t, gather_idx in packs:
            # any value near the column mean conditions the fp64 moments: the first row's (every rank holds the whole cohort and reduces its
            # share of the rows, so all ranks pick the same shift)
            shift = data[0, gather_idx].contiguous()
            parts.append(_moments_batched(data, ci_t, shift, rows))
        return D.all_reduce_sum_(torch.cat(parts))

    @staticmethod
    def _scores_from_moments(flat: np.ndarray, member_cols: List[List[int]]) -> List[float]:
        scores = []
        for i, cols in enumerate(member_cols):
            k = len(cols)
            blk = flat[i]
            mom = np.concatenate([blk[:1], blk[1:1 + k], blk[33:].reshape(32, 32)[:k, :k].reshape(-1)])
            corr = _corr_from_moments(mom, k)
            scores.append(float(corr[np.triu_indices(k, k=1)].mean()))
        return scores

    def _coherence_scores(self, data: torch.Tensor, member_cols: List[List[int]]) -> List[float]:
        return self._scores_from_moments(self._coherence_moments(data, member_cols).cpu().numpy(), member_cols)

    def _coherence_scores_pair(self, real: torch.Tensor, members_real, synthetic: torch.Tensor, members_syn):
        """Both cohorts enqueued back to back, ONE device-to-host copy for the pair."""
        mr = self._coherence_moments(real, members_real)
        ms = self._coherence_moments(synthetic, members_syn)
        flat = torch.cat([mr, ms]).cpu().numpy()
        return self._scores_from_moments(flat[:len(members_real)], members_real), self._scores_from_moments(flat[len(members_real):], members_syn)

    def validate_pathway_coherence(self, real_data, synthetic_data, pathway_gene_matrix) -> Dict[str, float]:
        """Mean within-pathway pairwise Pearson correlation for the first 10 pathways (>= 3 member genes present), for the
        real and the synthetic cohort, and the correlation of the two score vectors.
        real_data / synthetic_data: DataFrames with gene-symbol columns; pathway_gene_matrix: genes x pathways 0/1 DataFrame."""
        self._require_cuda()
        results: Dict[str, float] = {}
        real_cols = list(real_data.columns)
        syn_index = {g: i for i, g in enumerate(synthetic_data.columns)}
        real_index = {g: i for i, g in enumerate(real_cols)}
        members_real, members_syn = [], []
        for pathway in pathway_gene_matrix.columns[:10]:
            genes = pathway_gene_matrix[pathway_gene_matrix[pathway] == 1].index
            genes = [g for g in genes if g in real_index]
            if len(genes) < 3:
                continue
            if len(genes) > 32:
                raise ValueError("pathways with more than 32 member genes are not supported by the moment kernel")
            members_real.append([real_index[g] for g in genes])
            members_syn.append([syn_index[g] for g in genes])      # KeyError like the reference's synthetic_data[pathway_genes]
        if not members_real:
            return results
        real_t = _to_device(real_data, self.device)
        syn_t = _to_device(synthetic_data, self.device)
        real_scores, syn_scores = self._coherence_scores_pair(real_t, members_real, syn_t, members_syn)
        results["real_pathway_coherence"] = float(np.mean(real_scores))
        results["synthetic_pathway_coherence"] = float(np.mean(syn_scores))
        results["pathway_coherence_correlation"] = float(np.corrcoef(real_scores, syn_scores)[0, 1])
        logger.info("Real pathway coherence: %.3f", results["real_pathway_coherence"])
        logger.info("Synthetic pathway coherence: %.3f", results["synthetic_pathway_coherence"])
        logger.info("Coherence correlation: %.3f", results["pathway_coherence_correlation"])
        return results

    def pathway_coherence_from_tensors(self, real: torch.Tensor, synthetic: torch.Tensor, members: Sequence[Sequence[int]]) -> Dict[str, float]:
        """Tensor entry point for large GPU-resident cohorts: `members[p]` = column indices of pathway p's genes."""
        self._require_cuda()
        members = [list(m) for m in list(members)[:10] if len(m) >= 3]
        if not members:
            return {}
        rs, ss = self._coherence_scores_pair(_to_device(real, self.device), members, _to_device(synthetic, self.device), members)
        return {"real_pathway_coherence": float(np.mean(rs)), "synthetic_pathway_coherence": float(np.mean(ss)),
                "pathway_coherence_correlation": float(np.corrcoef(rs, ss)[0, 1])}

    # ------------------------------------------------------------------ utils/validation.py:177-223
    def validate_mutation_expression_correlation(self, mutations, expression, pathway_scores) -> Dict[str, float]:
        """Sign check of corr(mutation status, pathway activity) for every rule of evaluation.required_correlations."""
        self._require_cuda()
        results: Dict[str, float] = {}
        violations = total = 0
        rank, ws = D.world()
        for rule in self.required_correlations:
            gene, pathway, expected = rule["mutation"], rule["pathway"], rule["direction"]
            if gene not in mutations.columns or pathway not in pathway_scores.columns:
                continue
            pair = torch.stack([_to_device(mutations[gene], self.device), _to_device(pathway_scores[pathway], self.device)], dim=1).contiguous()
            rows = D.shard_rows(pair.shape[0], rank, ws)
            mom = D.all_reduce_sum_(_moments(pair, [0, 1], pair[0].contiguous(), rows)).cpu().numpy()
            corr = float(_corr_from_moments(mom, 2)[0, 1])
            if expected == "positive" and corr < 0:
                violations += 1
            elif expected == "negative" and corr > 0:
                violations += 1
            total += 1
            logger.info("%s vs %s: corr=%.3f (expected: %s)", gene, pathway, corr, expected)
        if total > 0:
            results["mutation_expression_violation_rate"] = violations / total
        return results
