"""Entity-ranking evaluation on the fused kernels: the caller on the far side of the scoring path.

Mirrors `EntityRankingJob._evaluate` (kge/job/eval_entity_ranking.py:103-487): for every evaluation triple the
true object is ranked among all objects of (s,p,?) and the true subject among all subjects of (?,p,o), raw and
filtered; metrics are mean rank, mean reciprocal rank and hits@k over both directions (:620-649), under the
reference's key names.  What differs is where the work happens:

 * true scores come from the sp_/_po path on the unique targets (:192-203), as in the reference;
 * scores are never materialised: every (ranking, direction, entity chunk) is ONE fused score+rank launch
   (`model.rank_sp / rank_po`), whose integer rank / tie counts add up over chunks (:310-313) — and over
   GPUs, see kge_b200.sharded;
 * filter labels come from the native KvsAllIndex in CSR form (kge_b200.indexing) and are densified per
   chunk on the device only for the batch at hand; "filtered_with_test" filters the test split's answers
   on top of the filtered ranking (:278-303), i.e. uses the sum of both filter matrices.

The `model` argument is duck-typed (kge_b200.KgeModel, the LibKGE plugin models, or a CPU stand-in in tests):
it needs `score_sp(s, p, o_subset)`, `score_po(p, o, s_subset)`, `rank_sp(s, p, true, entity_subset,
filter_labels, rtol, atol, rank, ties)` and `rank_po(p, o, true, ...)`.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Sequence

import torch

from . import indexing

S, P, O = 0, 1, 2


class EntityRankingEvaluator:
    def __init__(self, model, num_entities: int, filter_splits: Sequence[torch.Tensor],
                 test_triples: Optional[torch.Tensor] = None, batch_size: int = 100, chunk_size: int = -1,
                 tie_handling: str = "rounded_mean_rank", rtol: float = 1e-4, atol: float = 1e-5,
                 hits_at_k_s: Iterable[int] = (1, 3, 10, 50, 100, 200, 300, 400, 500, 1000), device=None,
                 warn_only: bool = False):
        if tie_handling not in ("rounded_mean_rank", "best_rank", "worst_rank"):
            raise NotImplementedError(tie_handling)       # eval_entity_ranking.py:616-618
        self.model, self.E = model, int(num_entities)
        self.batch_size, self.chunk_size = int(batch_size), int(chunk_size)
        self.tie_handling, self.rtol, self.atol = tie_handling, rtol, atol
        self.warn_only = warn_only                                            # entity_ranking.tie_handling.warn_only
        self.hits_at_k_s = [k for k in hits_at_k_s if k <= self.E] or [1]     # :45-52 (k capped by #entities)
        self.device = device
        known = torch.cat([t.view(-1, 3).long().cpu() for t in filter_splits], 0) if len(filter_splits) else \
            torch.zeros((0, 3), dtype=torch.long)
        self._sp = indexing.index_KvsAll(known, "sp")
        self._po = indexing.index_KvsAll(known, "po")
        self._test_sp = self._test_po = None
        if test_triples is not None:
            t = test_triples.view(-1, 3).long().cpu()
            self._test_sp, self._test_po = indexing.index_KvsAll(t, "sp"), indexing.index_KvsAll(t, "po")

    # -- filter labels -----------------------------------------------------------------------------------
    def _coords(self, batch_cpu, sp, po):
        offs, cols = indexing.sp_po_label_csr(batch_cpu, self.E, sp, po)
        return indexing.csr_to_coords(offs, cols)

    def _dense_chunk(self, coords, n, lo, hi, s, o, dev):
        """[2, n, hi-lo] with +inf at known answers inside the entity chunk, the example's own answer zeroed
        (:270-290).  Plane 0 filters score_sp (objects), plane 1 score_po (subjects)."""
        f = torch.zeros((2, n, hi - lo), dtype=torch.float32, device=dev)
        if coords.numel():
            c = coords.to(dev)
            plane = (c[:, 1] >= self.E).long()
            col = c[:, 1] - plane * self.E
            keep = (col >= lo) & (col < hi)
            f[plane[keep], c[keep, 0], col[keep] - lo] = float("inf")
        rows = torch.arange(n, device=dev)
        m = (o >= lo) & (o < hi)
        f[0, rows[m], o[m] - lo] = 0.0
        m = (s >= lo) & (s < hi)
        f[1, rows[m], s[m] - lo] = 0.0
        return f

    # -- one batch ---------------------------------------------------------------------------------------
    def _rank_batch(self, batch_cpu: torch.Tensor) -> Dict[str, torch.Tensor]:
        dev = self.device
        b = batch_cpu.to(dev) if dev is not None else batch_cpu
        dev = b.device
        s, p, o = b[:, S].long().contiguous(), b[:, P].long().contiguous(), b[:, O].long().contiguous()
        n = s.numel()
        uo, uo_inv = torch.unique(o, return_inverse=True)
        o_true = torch.gather(self.model.score_sp(s, p, uo), 1, uo_inv.view(-1, 1)).view(-1).contiguous()
        us, us_inv = torch.unique(s, return_inverse=True)
        s_true = torch.gather(self.model.score_po(p, o, us), 1, us_inv.view(-1, 1)).view(-1).contiguous()

        coords = self._coords(batch_cpu, self._sp, self._po)
        coords_test = self._coords(batch_cpu, self._test_sp, self._test_po) if self._test_sp is not None else None
        rankings = ["", "_filtered"] + (["_filtered_with_test"] if coords_test is not None else [])
        counts = {r: [torch.zeros(n, dtype=torch.int64, device=dev) for _ in range(4)] for r in rankings}
        chunk = self.E if self.chunk_size <= 0 else self.chunk_size
        for c in range(math.ceil(self.E / chunk)):
            lo, hi = c * chunk, min((c + 1) * chunk, self.E)
            subset = None if (lo == 0 and hi == self.E) else torch.arange(lo, hi, device=dev)
            f = self._dense_chunk(coords, n, lo, hi, s, o, dev)
            filt = {"": None, "_filtered": f}
            if coords_test is not None:
                filt["_filtered_with_test"] = f + self._dense_chunk(coords_test, n, lo, hi, s, o, dev)
            for r in rankings:
                s_rank, s_ties, o_rank, o_ties = counts[r]
                fl = filt[r]
                self.model.rank_sp(s, p, o_true, subset, None if fl is None else fl[0], self.rtol, self.atol,
                                   o_rank, o_ties)
                self.model.rank_po(p, o, s_true, subset, None if fl is None else fl[1], self.rtol, self.atol,
                                   s_rank, s_ties)
        out = {}
        for r in rankings:
            s_rank, s_ties, o_rank, o_ties = counts[r]
            # the true answer counts as a tie of itself, so ties >= 1 in every ranking; 0 means the fused rank
            # kernel's value at the true column left the tolerance band around the precomputed true score — the
            # condition the reference checks at eval_entity_ranking.py:240-274 ("Error in tie-handling")
            bad = int((s_ties < 1).sum()) + int((o_ties < 1).sum())
            if bad:
                msg = (f"Error in tie-handling: {bad} true answers of ranking '{r or '_raw'}' fall outside the "
                       f"tolerance band (rtol={self.rtol}, atol={self.atol}) around their true scores")
                if not self.warn_only:
                    raise ValueError(msg)
                import warnings
                warnings.warn(msg)
                s_ties.clamp_(min=1)
                o_ties.clamp_(min=1)
            out["s" + r] = self._final(s_rank, s_ties)
            out["o" + r] = self._final(o_rank, o_ties)
        return out

    def _final(self, rank, ties):
        if self.tie_handling == "rounded_mean_rank":      # :598-618
            return rank + ties // 2
        if self.tie_handling == "best_rank":
            return rank
        return rank + ties - 1

    # -- whole split -------------------------------------------------------------------------------------
    def evaluate(self, triples: torch.Tensor, return_ranks: bool = False):
        """Metrics over `triples` [N,3]; with return_ranks also the per-example 0-based ranks per ranking."""
        tri = triples.view(-1, 3).cpu()
        hists: Dict[str, torch.Tensor] = {}
        ranks: Dict[str, List[torch.Tensor]] = {}
        for i in range(0, tri.shape[0], self.batch_size):
            res = self._rank_batch(tri[i:i + self.batch_size])
            for key, r in res.items():
                suffix = key[1:]
                h = hists.setdefault(suffix, torch.zeros(self.E, dtype=torch.float64))
                h += torch.bincount(r.cpu(), minlength=self.E).double()[: self.E]     # hist_all :652-669
                if return_ranks:
                    ranks.setdefault(key, []).append(r.cpu())
        metrics: Dict[str, float] = {}
        for suffix, h in hists.items():
            metrics.update(self.compute_metrics(h, suffix))
        if return_ranks:
            return metrics, {k: torch.cat(v) for k, v in ranks.items()}
        return metrics

    def compute_metrics(self, rank_hist: torch.Tensor, suffix: str = "") -> Dict[str, float]:
        """_compute_metrics (:620-649) on a histogram of 0-based ranks."""
        n = float(rank_hist.sum())
        r1 = torch.arange(1, self.E + 1, dtype=torch.float64)
        m = {
            "mean_rank" + suffix: float((rank_hist * r1).sum() / n) if n > 0 else 0.0,
            "mean_reciprocal_rank" + suffix: float((rank_hist / r1).sum() / n) if n > 0 else 0.0,
        }
        cum = torch.cumsum(rank_hist[: max(self.hits_at_k_s)], 0)
        for k in self.hits_at_k_s:
            m[f"hits_at_{k}{suffix}"] = float(cum[k - 1] / n) if n > 0 else 0.0
        return m
