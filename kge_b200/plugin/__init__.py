"""LibKGE plugin: the reference's own model classes with the scoring path served by libb200kge.

Usage (unchanged LibKGE jobs, README.md:522-563 plugin mechanism):

    modules: [kge.job, kge.model, kge.model.embedder, kge_b200.plugin]
    model: b200_complex            # or b200_distmult / b200_simple / b200_cp / b200_rescal /
                                   #    b200_transe / b200_rotate

`b200_<model>.yaml` in this directory mirrors the reference's `<model>.yaml` (same embedder keys, so
`_entity_embedder._embeddings.weight` checkpoints load unchanged) and adds `precision`.
The classes subclass the reference's `KgeModel` / `RelationalScorer` (kge_model.py:122,354):

 * `score_emb(s_emb, p_emb, o_emb, combine)` (scorer level) → b200kge_score_spo / score_1vsN; this also
   serves `ReciprocalRelationsModel`, which calls the scorer directly
   (reciprocal_relations_model.py:76-124).
 * `score_spo / score_sp / score_po / score_sp_po` (model level) → index-level entry points that read
   the embedding tables in place (the LookupEmbedder gather is fused, no `embed_all()` table copy)
   whenever both embedders are plain LookupEmbedders with dropout inactive; otherwise the
   reference's embedders run and the scorer-level path takes over.
 * extra fused methods `score_sp_loss / score_po_loss / rank_sp / rank_po` for job plugins.

CUDA only: CPU tensors raise (no fallback).  Backward ("next") is provided by
recomputation through the reference's own dense expression inside a torch.autograd.Function.

This module needs the reference package `kge` to be importable; kge_b200's standalone mirror
(kge_b200.model) does not.
"""
from __future__ import annotations

import torch

from kge.model import ComplEx, CP, DistMult, Rescal, RotatE, SimplE, TransE
from kge.model.complex import ComplExScorer
from kge.model.cp import CPScorer
from kge.model.distmult import DistMultScorer
from kge.model.embedder.lookup_embedder import LookupEmbedder
from kge.model.rescal import RescalScorer
from kge.model.rotate import RotatEScorer
from kge.model.simple import SimplEScorer
from kge.model.transe import TransEScorer

from .. import engine


class _ScoreEmbFn(torch.autograd.Function):
    """Forward: sm_100a kernels.  Backward: recompute with the reference's dense expression."""

    @staticmethod
    def forward(ctx, scorer, ref_score_emb, combine, s_emb, p_emb, o_emb):
        ctx.ref, ctx.combine = ref_score_emb, combine
        ctx.save_for_backward(s_emb, p_emb, o_emb)
        return scorer._b200_forward(s_emb.detach(), p_emb.detach(), o_emb.detach(), combine)

    @staticmethod
    def backward(ctx, grad_out):
        s, p, o = (t.detach().requires_grad_(True) for t in ctx.saved_tensors)
        with torch.enable_grad():
            out = ctx.ref(s, p, o, ctx.combine)
            gs, gp, go = torch.autograd.grad(out, (s, p, o), grad_out.reshape(out.shape), allow_unused=True)
        return None, None, None, gs, gp, go


class _B200ScorerMixin:
    _b200_name = None

    def _b200_l_norm(self):
        return float(getattr(self, "_norm", 1.0))

    def _b200_precision(self):
        try:
            return self.get_option("precision")
        except Exception:
            return "auto"

    def _b200_forward(self, s_emb, p_emb, o_emb, combine):
        name, ln, prec = self._b200_name, self._b200_l_norm(), self._b200_precision()
        n = p_emb.size(0)
        if combine == "spo":
            return engine.score_spo(name, s_emb, p_emb, o_emb, l_norm=ln).view(n, -1)
        if combine == "sp_":
            return engine.score_1vsN(name, "sp_", s_emb, p_emb, o_emb, l_norm=ln, precision=prec)
        if combine == "_po":
            return engine.score_1vsN(name, "_po", o_emb, p_emb, s_emb, l_norm=ln, precision=prec)
        raise ValueError(combine)

    def score_emb(self, s_emb, p_emb, o_emb, combine: str):
        if combine not in ("spo", "sp_", "_po"):
            # "s_o" is outside the fused scope: generic expansion of the base class, which lands
            # in this class again with combine="spo" (kge_model.py:200-209)
            return super().score_emb(s_emb, p_emb, o_emb, combine)
        ref = super().score_emb
        needs_grad = torch.is_grad_enabled() and any(t.requires_grad for t in (s_emb, p_emb, o_emb))
        if needs_grad:
            return _ScoreEmbFn.apply(self, ref, combine, s_emb, p_emb, o_emb)
        return self._b200_forward(s_emb, p_emb, o_emb, combine)


def _scorer(name, base):
    return type(f"B200{base.__name__}", (_B200ScorerMixin, base), {"_b200_name": name})


class _B200ModelMixin:
    """Index-level overrides (kge_model.py:663-789): read the tables in place when possible."""

    _b200_name = None
    _b200_scorer_cls = None

    def __init__(self, config, dataset, configuration_key=None, init_for_load_only=False):
        super().__init__(config=config, dataset=dataset, configuration_key=configuration_key,
                         init_for_load_only=init_for_load_only)
        # swap the reference scorer for ours (same configuration key, same options)
        self._scorer = self._b200_scorer_cls(config, dataset, self.configuration_key)

    # -- helpers
    def _b200_direct(self):
        """True if the tables can be read in place: plain LookupEmbedders, dropout inactive, and no
        autograd graph requested (training backward goes through the scorer-level Function)."""
        es, ep, eo = self.get_s_embedder(), self.get_p_embedder(), self.get_o_embedder()
        for e in (es, ep, eo):
            if type(e) is not LookupEmbedder:
                return False
            if e.dropout.p > 0 and e.training:
                return False
        if es is not eo:
            return False
        w = es._embeddings.weight
        if torch.is_grad_enabled() and (w.requires_grad or ep._embeddings.weight.requires_grad):
            return False
        return True

    def _b200_tables(self):
        return (self.get_s_embedder()._embeddings.weight.detach(),
                self.get_p_embedder()._embeddings.weight.detach())

    def _b200_args(self):
        sc = self._scorer
        return sc._b200_l_norm(), sc._b200_precision()

    def score_spo(self, s, p, o, direction=None):
        if not self._b200_direct():
            return super().score_spo(s, p, o, direction)
        ent, rel = self._b200_tables()
        return engine.score_spo(self._b200_name, ent, rel, ent, s, p, o, self._b200_args()[0]).view(-1)

    def score_sp(self, s, p, o=None):
        if not self._b200_direct():
            return super().score_sp(s, p, o)
        ent, rel = self._b200_tables()
        ln, prec = self._b200_args()
        return engine.score_1vsN(self._b200_name, "sp_", ent, rel, ent, s, p, o, ln, prec)

    def score_po(self, p, o, s=None):
        if not self._b200_direct():
            return super().score_po(p, o, s)
        ent, rel = self._b200_tables()
        ln, prec = self._b200_args()
        return engine.score_1vsN(self._b200_name, "_po", ent, rel, ent, o, p, s, ln, prec)

    def score_sp_po(self, s, p, o, entity_subset=None):
        if not self._b200_direct():
            return super().score_sp_po(s, p, o, entity_subset)
        ent, rel = self._b200_tables()
        ln, prec = self._b200_args()
        return engine.score_sp_po(self._b200_name, ent, rel, s, p, o, entity_subset, ln, prec)

    # -- fused forms for job plugins (scores never reach HBM); forward only
    def score_sp_loss(self, s, p, labels, loss="bce", offset=0.0):
        ent, rel = self._b200_tables()
        ln, prec = self._b200_args()
        return engine.score_1vsN_loss(self._b200_name, "sp_", ent, rel, ent, labels, s, p, None, loss, offset, ln, prec)

    def score_po_loss(self, p, o, labels, loss="bce", offset=0.0):
        ent, rel = self._b200_tables()
        ln, prec = self._b200_args()
        return engine.score_1vsN_loss(self._b200_name, "_po", ent, rel, ent, labels, o, p, None, loss, offset, ln, prec)

    def rank_sp(self, s, p, true_scores, entity_subset=None, filter_labels=None, rtol=1e-4, atol=1e-5,
                rank=None, ties=None):
        ent, rel = self._b200_tables()
        ln, prec = self._b200_args()
        return engine.score_1vsN_rank(self._b200_name, "sp_", ent, rel, ent, true_scores, s, p, entity_subset,
                                      filter_labels, rtol, atol, ln, prec, rank, ties)

    def rank_po(self, p, o, true_scores, entity_subset=None, filter_labels=None, rtol=1e-4, atol=1e-5,
                rank=None, ties=None):
        ent, rel = self._b200_tables()
        ln, prec = self._b200_args()
        return engine.score_1vsN_rank(self._b200_name, "_po", ent, rel, ent, true_scores, o, p, entity_subset,
                                      filter_labels, rtol, atol, ln, prec, rank, ties)


def _model(cls_name, name, base_model, base_scorer):
    return type(cls_name, (_B200ModelMixin, base_model),
                {"_b200_name": name, "_b200_scorer_cls": _scorer(name, base_scorer)})


B200ComplEx = _model("B200ComplEx", "complex", ComplEx, ComplExScorer)
B200DistMult = _model("B200DistMult", "distmult", DistMult, DistMultScorer)
B200SimplE = _model("B200SimplE", "simple", SimplE, SimplEScorer)
B200CP = _model("B200CP", "cp", CP, CPScorer)
B200Rescal = _model("B200Rescal", "rescal", Rescal, RescalScorer)
B200TransE = _model("B200TransE", "transe", TransE, TransEScorer)
B200RotatE = _model("B200RotatE", "rotate", RotatE, RotatEScorer)

__all__ = ["B200ComplEx", "B200DistMult", "B200SimplE", "B200CP", "B200Rescal", "B200TransE", "B200RotatE"]


def install_native_indexes(dataset, splits=("train", "valid", "test")):
    """Serve the dataset's `{split}_{sp|po|so}_to_{o|s|p}` indexes (kge/indexing.py:197-235) from
    kge_b200.indexing.KvsAllIndex — same attributes and accessors as the reference class (TrainingJobKvsAll's
    collate and EntityRankingJob's label lookup use it unchanged), built by the native sort/unique/lookup code
    instead of numpy + a numba dict.  Call once after the dataset is created."""
    from ..indexing import KvsAllIndex

    def make(split, key, cols, val, name):
        def fn(ds):
            if not ds._indexes.get(name):
                ds._indexes[name] = KvsAllIndex(ds.split(split), cols, val, torch.IntTensor)
            ds.config.log("{} distinct {} pairs in {}".format(len(ds._indexes[name]), key, split), prefix="  ")
            return ds._indexes.get(name)
        return fn

    for split in splits:
        for key, cols, val, v in (("sp", [0, 1], 2, "o"), ("po", [1, 2], 0, "s"), ("so", [0, 2], 1, "p")):
            name = f"{split}_{key}_to_{v}"
            dataset.index_functions[name] = make(split, key, cols, val, name)
            dataset._indexes.pop(name, None)
    return dataset
