"""GPU mirror of `BiologicalValidator` (utils/validation.py:18-387): the three validators on the hot path --
`compute_mmd`, `validate_pathway_coherence`, `validate_mutation_expression_correlation` (:273-298, :125-175, :177-223) --
and the rest of the class (`validate_mutation_cooccurrence` :27-123, `statistical_tests` :225-271, `validate_all` :300-387;
SURVEY.md §8(f)#3) so that `main.py --steps validate` (main.py:322) runs against the drop-in.  Same method names, argument
meaning, result keys and return values as the reference.

The heavy arithmetic runs in the C-ABI library: RBF Gram tiles on tcgen05 with an exp/sum epilogue; column-gathered fp64
moment reduction for the Pearson correlations; the 2x2 co-occurrence tables as ONE binary M^T M contraction on tcgen05
(`osteo_wgrad_tc`: 0/1 is exact in bf16 and the fp32 accumulators count exactly up to 2^24 rows); the PCA scatter matrix as a
split-bf16 (fp32x3) Gram.  What is left to torch on the device are library primitives outside the hot path: `sort` /
`searchsorted` for the KS statistic and the 1-D Wasserstein distance, `linalg.eigh` for the 10 leading principal axes.  Scalar
finishes (Pearson from moments, chi-square from four counts, KS p-value from D) run on the host in float64 like the reference.
Under torch.distributed the Gram rows / cohort rows are sharded over ranks and only the partial sums are all-reduced
(SURVEY.md §8e).  No CPU fallback.
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


def _gram_partial_sums_cyclic(X: torch.Tensor, Y: torch.Tensor, gamma: float, center: torch.Tensor, rank: int, world: int, precision: int) -> torch.Tensor:
    """This rank's share {sum Kxx, sum Kyy, sum Kxy} under block-cyclic Gram-row sharding (symmetric half-Grams for Kxx / Kyy)."""
    sums = torch.zeros(3, dtype=torch.float64, device=X.device)
    _lib.check(_lib.load().osteo_mmd_partial_cyclic(X.data_ptr(), X.shape[0], Y.data_ptr(), Y.shape[0], X.shape[1], float(gamma), center.data_ptr(),
                                                    int(rank), int(world), precision, sums.data_ptr(), _lib.stream_handle()))
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


def _moments_tiled(data: torch.Tensor, ci_t: torch.Tensor, rows, max_set_size: int = 32, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp64 [n_sets, _CM_STRIDE] like _moments_batched by the register-tiled kernel (osteo_corr_moments_tiled); the shift of a column is
    its value in row 0 of `data` (taken inside the kernel). max_set_size: the largest set (<= 16 selects the 10-block form).
    `out`: an existing contiguous [n_sets, _CM_STRIDE] fp64 view to fill (no allocation, no later concatenation)."""
    if out is None:
        out = torch.empty((ci_t.shape[0], _CM_STRIDE), dtype=torch.float64, device=data.device)
    _lib.check(_lib.load().osteo_corr_moments_tiled(data.data_ptr(), data.shape[0], data.stride(0), data.shape[1], ci_t.data_ptr(), ci_t.shape[0],
                                                    int(max_set_size), None, rows[0], rows[1], out.data_ptr(), _lib.stream_handle()))
    return out


def _coherence_finish(mom: torch.Tensor, ci_t: torch.Tensor) -> torch.Tensor:
    """fp64 [n_sets] device tensor: mean upper-triangle Pearson correlation per set from (all-reduced) moment blocks."""
    scores = torch.empty(mom.shape[0], dtype=torch.float64, device=mom.device)
    _lib.check(_lib.load().osteo_coherence_finish(mom.data_ptr(), ci_t.data_ptr(), mom.shape[0], scores.data_ptr(), _lib.stream_handle()))
    return scores


def _corr_from_moments(mom: np.ndarray, k: int) -> np.ndarray:
    """Pearson correlation matrix from shifted moments (float64, host; k <= 32)."""
    n = mom[0]
    s1 = mom[1:1 + k]
    s2 = mom[1 + k:].reshape(k, k)
    cov = s2 - np.outer(s1, s1) / n
    sd = np.sqrt(np.diag(cov))
    with np.errstate(divide="ignore", invalid="ignore"):
        return cov / np.outer(sd, sd)


_on_validator_device = _lib.on_device(lambda v: v.device)


class BiologicalValidator:
    """Validate synthetic patients against biological knowledge (GPU-resident)."""

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
    @_on_validator_device
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
        # identical cohorts: the reference's three Grams are then the same float64 numbers and cancel to exactly 0.0; here K(X,X) is a
        # symmetric half-Gram and K(X,Y) a full one, whose fp32 tile sums would leave sqrt(1e-8). One O(n d) comparison settles it.
        if n == m and (X.data_ptr() == Y.data_ptr() or bool(torch.equal(X, Y))):
            return 0.0
        # RBF is translation invariant: centre on the pooled mean before the bf16 split (SURVEY.md §7 "MMD precision")
        center = ((X.sum(0, dtype=torch.float64) + Y.sum(0, dtype=torch.float64)) / (n + m)).float().contiguous()
        rank, ws = D.world()
        if ws > 1:
            # Gram rows block-cyclic over the ranks, Kxx / Kyy as symmetric half-Grams: balanced, and 2/3 of the tiles of full Grams
            sums = _gram_partial_sums_cyclic(X, Y, gamma, center, rank, ws, _PRECISIONS[self.precision])
        else:
            sums = _gram_partial_sums(X, Y, gamma, center, (0, n), (0, m), _PRECISIONS[self.precision])
        D.all_reduce_sum_(sums)
        sxx, syy, sxy = (float(v) for v in sums.cpu())
        mmd = sxx / (float(n) * n) + syy / (float(m) * m) - 2.0 * sxy / (float(n) * m)
        return float(np.sqrt(max(mmd, 0.0)))

    # ------------------------------------------------------------------ utils/validation.py:125-175
    def _index_packs(self, device, member_cols: List[List[int]]):
        """Cached device index tensors of the column sets: [(int32 [<= 32 sets, 32], -1 padded), ...]."""
        if any(len(c) > 32 for c in member_cols):
            raise ValueError("a pathway with more than 32 member genes is not supported by the moment kernel")
        key = (str(device), tuple(tuple(c) for c in member_cols))
        packs = self._index_cache.get(key)
        if packs is None:
            packs = []
            for b0 in range(0, len(member_cols), 32):
                sets = member_cols[b0:b0 + 32]
                ci = np.full((len(sets), 32), -1, dtype=np.int32)
                for i, cols in enumerate(sets):
                    ci[i, :len(cols)] = cols
                ci_t = torch.from_numpy(ci).to(device)
                packs.append((ci_t, ci_t.clamp(min=0).long()))
            if len(self._index_cache) > 16:
                self._index_cache.clear()
            self._index_cache[key] = packs
        return packs

    def _coherence_moments(self, data: torch.Tensor, member_cols: List[List[int]], out: Optional[torch.Tensor] = None, reduce: bool = True) -> torch.Tensor:
        """Device tensor [n_sets, _CM_STRIDE] of moment blocks (all-reduced over the ranks unless reduce=False): all pathways in one pass
        over the cohort by the register-tiled kernel (osteo_corr_moments_tiled). Nothing here waits for the device: the index tensors
        are cached per (device, column sets) and the result stays in HBM. The shift that conditions the fp64 moments is row 0 of the
        cohort, read inside the kernel (every rank holds the whole cohort and reduces its share of the rows: same shift everywhere)."""
        rank, ws = D.world()
        rows = D.shard_rows(data.shape[0], rank, ws)
        packs = self._index_packs(data.device, member_cols)
        kmax = max(len(c) for c in member_cols)
        if out is None:
            out = torch.empty((len(member_cols), _CM_STRIDE), dtype=torch.float64, device=data.device)
        s0 = 0
        for ci_t, _ in packs:
            _moments_tiled(data, ci_t, rows, kmax, out=out[s0:s0 + ci_t.shape[0]])
            s0 += ci_t.shape[0]
        return D.all_reduce_sum_(out) if reduce else out

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

    def _index_tensor(self, device, member_cols: List[List[int]]) -> torch.Tensor:
        packs = self._index_packs(device, member_cols)
        return packs[0][0] if len(packs) == 1 else torch.cat([ci for ci, _ in packs])

    def _coherence_scores_pair(self, real: torch.Tensor, members_real, synthetic: torch.Tensor, members_syn):
        """Both cohorts enqueued back to back into ONE moment buffer -- two moment kernels, one all-reduce (if sharded), one Pearson
        finish on the device -- and ONE device-to-host copy of the float64 scores of the pair."""
        nr, ns = len(members_real), len(members_syn)
        key = ("pair", str(real.device), tuple(tuple(c) for c in members_real), tuple(tuple(c) for c in members_syn))
        ci_all = self._index_cache.get(key)
        if ci_all is None:
            ci_all = torch.cat([self._index_tensor(real.device, members_real), self._index_tensor(synthetic.device, members_syn)])
            self._index_cache[key] = ci_all
        mom = torch.empty((nr + ns, _CM_STRIDE), dtype=torch.float64, device=real.device)
        self._coherence_moments(real, members_real, out=mom[:nr], reduce=False)
        self._coherence_moments(synthetic, members_syn, out=mom[nr:], reduce=False)
        D.all_reduce_sum_(mom)
        flat = _coherence_finish(mom, ci_all).cpu().numpy()
        return [float(v) for v in flat[:nr]], [float(v) for v in flat[nr:]]

    @_on_validator_device
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

    @_on_validator_device
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
    @_on_validator_device
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

    # ------------------------------------------------------------------ utils/validation.py:27-123
    def _binary_columns(self, df, cols: Sequence) -> torch.Tensor:
        """fp32 device tensor [n, len(cols)] of the given DataFrame columns; raises unless every value is 0 or 1 (the co-occurrence
        tables are counted by a binary Gram; the reference's crosstab would grow extra categories for other values)."""
        t = _to_device(df[list(cols)], self.device)
        if not bool(((t == 0) | (t == 1)).all()):
            raise ValueError("mutation matrices must be binary (0 / 1)")
        return t

    def _cooccurrence_counts(self, m: torch.Tensor) -> np.ndarray:
        """n11[i, j] = #rows with columns i and j both 1: M^T M on tcgen05 (exact for < 2^24 rows per call)."""
        n, k = m.shape
        out = torch.zeros((k, k), dtype=torch.float64, device=m.device)
        lib = _lib.load()
        step = 1 << 23
        for r0 in range(0, n, step):
            blk = m[r0:r0 + step].contiguous()
            g = torch.empty((k, k), dtype=torch.float32, device=m.device)
            _lib.check(lib.osteo_wgrad_tc(blk.data_ptr(), blk.data_ptr(), g.data_ptr(), blk.shape[0], k, k, _lib.PREC_BF16, _lib.stream_handle()))
            out += g.double()
        return out.cpu().numpy()

    @staticmethod
    def _chi2_2x2(n11: float, s_i: float, s_j: float, n: float) -> float:
        """scipy.stats.chi2_contingency(pd.crosstab(a, b))[0] for two 0/1 columns from their counts: a constant column makes the
        crosstab 1 x 2 (dof 0 -> 0.0); otherwise Pearson's statistic with Yates' continuity correction (dof 1)."""
        if s_i in (0.0, n) or s_j in (0.0, n):
            return 0.0
        obs = np.array([[n - s_i - s_j + n11, s_j - n11], [s_i - n11, n11]], dtype=np.float64)
        exp = np.outer(obs.sum(1), obs.sum(0)) / n
        diff = exp - obs
        obs = obs + np.minimum(0.5, np.abs(diff)) * np.sign(diff)
        return float(((obs - exp) ** 2 / exp).sum())

    @_on_validator_device
    def validate_mutation_cooccurrence(self, real_mutations, synthetic_mutations) -> Dict[str, float]:
        """Mutation-frequency correlation, driver-gene frequency difference, mutual-exclusivity violation rate and the correlation
        of the pairwise chi-square scores of (up to) 50 randomly chosen genes -- drawn with the same `np.random.choice` call as the
        reference, so a seeded numpy RNG picks the same genes."""
        self._require_cuda()
        results: Dict[str, float] = {}
        real_t = _to_device(real_mutations, self.device)
        syn_t = _to_device(synthetic_mutations, self.device)
        real_cols, syn_cols = list(real_mutations.columns), list(synthetic_mutations.columns)
        real_freq = (real_t.sum(0, dtype=torch.float64) / real_t.shape[0]).cpu().numpy()
        syn_freq = (syn_t.sum(0, dtype=torch.float64) / syn_t.shape[0]).cpu().numpy()
        ri = {g: i for i, g in enumerate(real_cols)}
        si = {g: i for i, g in enumerate(syn_cols)}
        common_genes = real_mutations.columns.intersection(synthetic_mutations.columns)
        common = list(common_genes)
        freq_corr = np.corrcoef(real_freq[[ri[g] for g in common]], syn_freq[[si[g] for g in common]])[0, 1]
        results["mutation_frequency_correlation"] = freq_corr
        logger.info("Mutation frequency correlation: %.3f", freq_corr)

        drivers = [g for g in self.driver_genes if g in ri]
        if drivers:
            diff = np.abs(real_freq[[ri[g] for g in drivers]] - syn_freq[[si[g] for g in drivers]]).mean()      # KeyError like the reference
            results["driver_gene_frequency_diff"] = diff
            logger.info("Driver gene frequency difference: %.3f", diff)

        if self.mutually_exclusive_pairs:
            violations, total_pairs = 0, 0
            for g1, g2 in self.mutually_exclusive_pairs:
                if g1 in si and g2 in si:
                    violations += int(((syn_t[:, si[g1]] == 1) & (syn_t[:, si[g2]] == 1)).sum())
                    total_pairs += 1
            if total_pairs > 0:
                results["mutual_exclusivity_violation_rate"] = violations / (syn_t.shape[0] * total_pairs)
                logger.info("Mutual exclusivity violation rate: %.3f", results["mutual_exclusivity_violation_rate"])

        sample_genes = np.random.choice(common_genes, size=min(50, len(common_genes)), replace=False)
        if len(sample_genes) >= 2:
            chi = []
            for t, index in ((real_t, ri), (syn_t, si)):
                m = t[:, [index[g] for g in sample_genes]].contiguous()
                if not bool(((m == 0) | (m == 1)).all()):
                    raise ValueError("mutation matrices must be binary (0 / 1)")
                n11 = self._cooccurrence_counts(m)
                s, n = np.diag(n11), float(m.shape[0])
                k = len(sample_genes)
                chi.append([self._chi2_2x2(n11[i, j], s[i], s[j], n) for i in range(k) for j in range(i + 1, k)])
            chi2_corr = np.corrcoef(chi[0], chi[1])[0, 1]
            results["cooccurrence_pattern_correlation"] = chi2_corr
            logger.info("Co-occurrence pattern correlation: %.3f", chi2_corr)
        return results

    # ------------------------------------------------------------------ utils/validation.py:225-271
    @staticmethod
    def _ks_pvalue(d: float, n1: int, n2: int) -> float:
        """Two-sided p-value of scipy.stats.ks_2samp(method='auto') from the statistic: exact for max(n1, n2) <= 10 000 (scipy's own
        lattice-path count, a scalar routine), asymptotic Kolmogorov distribution otherwise."""
        from scipy.stats import distributions

        if max(n1, n2) <= 10000:
            try:
                from scipy.stats._stats_py import _attempt_exact_2kssamp

                ok, _, prob = _attempt_exact_2kssamp(n1, n2, int(np.gcd(n1, n2)), d, "two-sided")
                if ok:
                    return float(np.clip(prob, 0, 1))
            except ImportError:      # private helper moved: fall through to the asymptotic form
                pass
        m, n = sorted([float(n1), float(n2)], reverse=True)
        return float(np.clip(distributions.kstwo.sf(d, np.round(m * n / (m + n))), 0, 1))

    @staticmethod
    def _cdf_pair(a: torch.Tensor, b: torch.Tensor):
        """Column-wise empirical CDFs of a [n1, k] and b [n2, k] evaluated at the pooled sorted sample: (all_sorted [n1+n2, k], F_a, F_b)."""
        a_s = torch.sort(a.t().contiguous(), dim=1).values
        b_s = torch.sort(b.t().contiguous(), dim=1).values
        allv = torch.sort(torch.cat([a_s, b_s], dim=1), dim=1).values
        fa = torch.searchsorted(a_s, allv, right=True).double() / a.shape[0]
        fb = torch.searchsorted(b_s, allv, right=True).double() / b.shape[0]
        return allv, fa, fb

    def _ks_statistics(self, a: torch.Tensor, b: torch.Tensor) -> np.ndarray:
        _, fa, fb = self._cdf_pair(a, b)
        d = fa - fb
        return torch.maximum(torch.clamp(-d.min(dim=1).values, 0, 1), d.max(dim=1).values).cpu().numpy()

    def _wasserstein(self, a: torch.Tensor, b: torch.Tensor) -> np.ndarray:
        """scipy.stats.wasserstein_distance per column: integral of |F_a - F_b| over the pooled sample."""
        allv, fa, fb = self._cdf_pair(a.double(), b.double())
        return ((fa - fb).abs()[:, :-1] * (allv[:, 1:] - allv[:, :-1])).sum(dim=1).cpu().numpy()

    # tcgen05 contractions behind the PCA (split-bf16 = fp32x3: ~3e-6 relative per entry)
    @staticmethod
    def _gram_rows(x: torch.Tensor) -> torch.Tensor:
        """x x^T, [n, n] fp32."""
        n, d = x.shape
        g = torch.empty((n, n), dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().osteo_linear_tc(x.data_ptr(), x.data_ptr(), None, g.data_ptr(), n, n, d, _lib.PREC_FP32X3, _lib.stream_handle()))
        return g

    @staticmethod
    def _gram_cols(x: torch.Tensor) -> torch.Tensor:
        """x^T x, [d, d] fp32 (both operands MN-major, rows split over CTAs)."""
        n, d = x.shape
        g = torch.empty((d, d), dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().osteo_wgrad_tc(x.data_ptr(), x.data_ptr(), g.data_ptr(), n, d, d, _lib.PREC_FP32X3, _lib.stream_handle()))
        return g

    @staticmethod
    def _project(x: torch.Tensor, axes_t: torch.Tensor) -> torch.Tensor:
        """x axes_t^T, [n, k] fp32."""
        n, d = x.shape
        k = axes_t.shape[0]
        out = torch.empty((n, k), dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().osteo_linear_tc(x.data_ptr(), axes_t.data_ptr(), None, out.data_ptr(), n, k, d, _lib.PREC_FP32X3, _lib.stream_handle()))
        return out

    def _pca_scores(self, real: torch.Tensor, synthetic: torch.Tensor, k: int = 10):
        """Scores of both cohorts on the k leading principal axes of the REAL cohort (sklearn PCA.fit_transform / .transform: centre on
        the real mean, project). The scatter matrix is a split-bf16 (fp32x3) Gram on tcgen05 -- rows x rows when there are fewer
        patients than features, features x features otherwise -- its k leading eigenvectors come from linalg.eigh in float64."""
        n, d = real.shape
        mean = real.sum(0, dtype=torch.float64) / n
        rc = (real.double() - mean).float().contiguous()
        sc = (synthetic.double() - mean).float().contiguous()
        k = min(k, n, d)
        if n <= d:
            g = self._gram_rows(rc).double()
            w, u = torch.linalg.eigh((g + g.t()) / 2)
            w, u = w[-k:].flip(0), u[:, -k:].flip(1)
            axes = (rc.double().t() @ u) / torch.sqrt(w.clamp_min(1e-300))          # [d, k] unit vectors
        else:
            g = self._gram_cols(rc).double()
            _, v = torch.linalg.eigh((g + g.t()) / 2)
            axes = v[:, -k:].flip(1)
        axes_t = axes.t().float().contiguous()                                        # [k, d]
        return self._project(rc, axes_t), self._project(sc, axes_t)

    @_on_validator_device
    def statistical_tests(self, real_data, synthetic_data) -> Dict[str, float]:
        """KS test on the first 100 features, RBF-MMD, mean 1-D Wasserstein distance over the 10 leading principal components."""
        self._require_cuda()
        results: Dict[str, float] = {}
        real = _to_device(real_data, self.device)
        syn = _to_device(synthetic_data, self.device)
        nf = min(real.shape[1], 100)
        d = self._ks_statistics(real[:, :nf], syn[:, :nf])
        pvals = np.array([self._ks_pvalue(float(x), real.shape[0], syn.shape[0]) for x in d])
        results["ks_test_mean_pvalue"] = np.mean(pvals)
        results["ks_test_fraction_significant"] = (pvals < 0.05).mean()
        logger.info("KS test mean p-value: %.3f", results["ks_test_mean_pvalue"])
        logger.info("KS test fraction significant: %.3f", results["ks_test_fraction_significant"])
        results["mmd"] = self.compute_mmd(real, syn)
        logger.info("MMD: %.4f", results["mmd"])
        real_pca, syn_pca = self._pca_scores(real, syn, 10)
        results["wasserstein_distance_mean"] = np.mean(self._wasserstein(real_pca, syn_pca))
        logger.info("Mean Wasserstein distance: %.3f", results["wasserstein_distance_mean"])
        return results

    # ------------------------------------------------------------------ utils/validation.py:300-387
    @_on_validator_device
    def validate_all(self, real_mutations, real_expression, real_pathways, synth_mutations, synth_expression, synth_pathways,
                     pathway_gene_matrix=None) -> Dict[str, float]:
        """Run all validation tests; same result keys and overall score as the reference."""
        all_results: Dict[str, float] = {}
        all_results.update(self.validate_mutation_cooccurrence(real_mutations, synth_mutations))
        if pathway_gene_matrix is not None:
            all_results.update(self.validate_pathway_coherence(real_expression, synth_expression, pathway_gene_matrix))
        all_results.update(self.validate_mutation_expression_correlation(synth_mutations, synth_expression, synth_pathways))
        real_all = torch.cat([_to_device(x, self.device) for x in (real_mutations, real_expression, real_pathways)], dim=1)
        syn_all = torch.cat([_to_device(x, self.device) for x in (synth_mutations, synth_expression, synth_pathways)], dim=1)
        all_results.update(self.statistical_tests(real_all, syn_all))
        for key, value in all_results.items():
            logger.info("%s: %.4f", key, value)
        score = []
        if "mutation_frequency_correlation" in all_results:
            score.append(all_results["mutation_frequency_correlation"])
        if "cooccurrence_pattern_correlation" in all_results:
            score.append(all_results["cooccurrence_pattern_correlation"])
        if "mutual_exclusivity_violation_rate" in all_results:
            score.append(1 - all_results["mutual_exclusivity_violation_rate"])
        if "mutation_expression_violation_rate" in all_results:
            score.append(1 - all_results["mutation_expression_violation_rate"])
        if score:
            all_results["overall_biological_score"] = np.mean(score)
            logger.info("Overall Biological Score: %.3f", all_results["overall_biological_score"])
        return all_results
