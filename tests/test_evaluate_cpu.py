"""Host logic of kge_b200.evaluate.EntityRankingEvaluator (batching, entity chunks, CSR filters from the native
index, filtered_with_test stacking, tie handling, metrics) against traces of the reference's own
EntityRankingJob (tests/golden/jobs_*.npz).  The model is a CPU stand-in with the oracle's arithmetic, so no
GPU is needed; on the GPU the same evaluator drives kge_b200.KgeModel (tests/test_gpu_model.py)."""
import os

import numpy as np
import pytest
import torch

from kge_b200.evaluate import EntityRankingEvaluator
from oracle import kge_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


class OracleModel:
    """score_sp / score_po / rank_sp / rank_po with the oracle's arithmetic (CPU)."""

    def __init__(self, model, ent, rel):
        self.m, self.ent, self.rel = model, ent, rel

    def score_sp(self, s, p, o=None):
        return orc.score_sp(self.m, self.ent, self.rel, s, p, o)

    def score_po(self, p, o, s=None):
        return orc.score_po(self.m, self.ent, self.rel, p, o, s)

    def _rank(self, scores, true, filt, rtol, atol, rank, ties):
        if filt is not None:
            scores = scores - filt
        r, t = orc.ranks_and_ties(scores, true, rtol, atol)
        rank += r
        ties += t
        return rank, ties

    def rank_sp(self, s, p, true, entity_subset=None, filter_labels=None, rtol=1e-4, atol=1e-5, rank=None, ties=None):
        return self._rank(self.score_sp(s, p, entity_subset), true, filter_labels, rtol, atol, rank, ties)

    def rank_po(self, p, o, true, entity_subset=None, filter_labels=None, rtol=1e-4, atol=1e-5, rank=None, ties=None):
        return self._rank(self.score_po(p, o, entity_subset), true, filter_labels, rtol, atol, rank, ties)


def _load(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(z[k]) if z[k].ndim else z[k].item() for k in z.files}


@pytest.mark.parametrize("model", ["complex", "transe"])
@pytest.mark.parametrize("batch_size,chunk_size", [(8, -1), (100, 7), (13, 20)])
def test_metrics_match_reference_job(model, batch_size, chunk_size):
    g = _load(f"jobs_{model}.npz")
    E = g["ent"].shape[0]
    ev = EntityRankingEvaluator(OracleModel(model, g["ent"], g["rel"]), E, [g["train"], g["valid"]], g["test"],
                                batch_size=batch_size, chunk_size=chunk_size, hits_at_k_s=(1, 3, 10, 50))
    assert ev.hits_at_k_s == [1, 3, 10]              # k capped by the number of entities
    met, ranks = ev.evaluate(g["valid"], return_ranks=True)
    for suffix in ("", "_filtered", "_filtered_with_test"):
        for k in ("mean_rank", "mean_reciprocal_rank", "hits_at_1", "hits_at_3", "hits_at_10"):
            want = float(g["valid_" + k + suffix])
            assert met[k + suffix] == pytest.approx(want, rel=1e-6, abs=1e-9), (k + suffix, met[k + suffix], want)
    n = g["valid"].shape[0]
    assert set(ranks) == {"s", "o", "s_filtered", "o_filtered", "s_filtered_with_test", "o_filtered_with_test"}
    assert all(r.shape == (n,) for r in ranks.values())
    # filtering can only improve a rank
    assert bool((ranks["o_filtered"] <= ranks["o"]).all()) and bool((ranks["s_filtered_with_test"] <= ranks["s_filtered"]).all())


def test_tie_handling_and_empty_split():
    g = _load("jobs_complex.npz")
    E = g["ent"].shape[0]
    ent = torch.zeros_like(g["ent"])                 # all scores equal: every entity ties with the answer
    m = OracleModel("complex", ent, g["rel"])
    n = 2 * g["valid"].shape[0]
    for th, mean_rank in (("best_rank", 1.0), ("worst_rank", float(E)), ("rounded_mean_rank", 1.0 + E // 2)):
        ev = EntityRankingEvaluator(m, E, [], tie_handling=th)
        met = ev.evaluate(g["valid"])
        assert met["mean_rank"] == pytest.approx(mean_rank)
    with pytest.raises(NotImplementedError):
        EntityRankingEvaluator(m, E, [], tie_handling="random")
    ev = EntityRankingEvaluator(m, E, [g["train"]])
    assert ev.evaluate(torch.zeros((0, 3), dtype=torch.long)) == {}
    assert ev.compute_metrics(torch.zeros(E, dtype=torch.float64))["mean_rank"] == 0.0
