"""CPU oracle for the LibKGE scoring hot path (TEST INFRASTRUCTURE — not product code).

This module restates, on the CPU with plain torch/numpy tensor expressions, the
arithmetic that the reference (uma-pi1/kge, "LibKGE") performs on the path

    LookupEmbedder gather -> RelationalScorer.score_emb -> KgeLoss / rank counting

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  The product package
``kge_b200`` never imports it and has no CPU fallback.

Parity pin: the reference holds no golden vectors for this path (its tests only check
self-consistency, tests/test_model.py:29-71).  The oracle is therefore pinned against
outputs of the reference itself, imported in the build container by
``tests/golden/gen_golden.py`` and committed as ``tests/golden/*.npz``
(``tests/test_oracle_golden.py`` replays them on every CPU test run).

Every function cites the reference file:line (relative to /root/reference) it follows.
All functions take/return torch CPU tensors; ``dtype`` may be float32 (what the
reference computes in) or float64 (ground truth for error budgets).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

S, P, O = 0, 1, 2

GEMM_FAMILY = ("complex", "distmult", "simple", "cp", "rescal")
DISTANCE_FAMILY = ("transe", "rotate")
MODELS = GEMM_FAMILY + DISTANCE_FAMILY


def relation_dim(model: str, dim: int) -> int:
    """Relation-embedding width for an entity width `dim`.

    cp.py:49-54 (D/2), rotate.py:92-97 (D/2 phases), rescal.py:78-95 (d*d), else D.
    """
    if model in ("cp", "rotate"):
        return dim // 2
    if model == "rescal":
        return dim * dim
    return dim


# --------------------------------------------------------------------------- a1 / a2
def embed(weight: torch.Tensor, indexes: torch.Tensor) -> torch.Tensor:
    """Row gather. lookup_embedder.py:96-97 (dropout p=0 => _postprocess is identity)."""
    return weight[indexes.long()]


def embed_all(weight: torch.Tensor) -> torch.Tensor:
    """Gather of arange(vocab): a fresh copy of the table. lookup_embedder.py:99-112."""
    return weight[torch.arange(weight.shape[0])]


# --------------------------------------------------------------------------- a4 .. a11
def _generic_expand(s, p, o, combine):
    """RelationalScorer.score_emb fallback: blow up to row-wise triples.

    kge_model.py:183-209.
    """
    if combine == "sp_":
        n, m = p.shape[0], o.shape[0]
        return s.repeat_interleave(m, 0), p.repeat_interleave(m, 0), o.repeat(n, 1), n
    if combine == "_po":
        n, m = p.shape[0], s.shape[0]
        return s.repeat(n, 1), p.repeat_interleave(m, 0), o.repeat_interleave(m, 0), n
    if combine == "s_o":
        n, m = s.shape[0], p.shape[0]
        return s.repeat_interleave(m, 0), p.repeat(n, 1), o.repeat_interleave(m, 0), n
    raise ValueError('cannot handle combine="{}"'.format(combine))


def _complex(s, p, o, combine):
    """complex.py:18-43 — four-block Hadamard form, GEMM over the 2D-wide operands."""
    h = p.shape[1] // 2
    p_re, p_im = p[:, :h], p[:, h:]
    o_re, o_im = o[:, :h], o[:, h:]
    s4 = torch.cat([s, s], 1)  # re im re im
    r4 = torch.cat([p_re, p_re, p_im, -p_im], 1)  # re re im -im
    o4 = torch.cat([o_re, o_im, o_im, o_re], 1)  # re im im re
    if combine == "spo":
        return (s4 * o4 * r4).sum(1)
    if combine == "sp_":
        return (s4 * r4) @ o4.t()
    if combine == "_po":
        return (r4 * o4) @ s4.t()
    return None


def _distmult(s, p, o, combine):
    """distmult.py:13-25."""
    if combine == "spo":
        return (s * p * o).sum(1)
    if combine == "sp_":
        return (s * p) @ o.t()
    if combine == "_po":
        return (o * p) @ s.t()
    return None


def _simple(s, p, o, combine):
    """simple.py:13-33 — head/tail halves, forward/backward relation halves, mean of two."""
    h = s.shape[1] // 2
    s_h, s_t = s[:, :h], s[:, h:]
    p_f, p_b = p[:, :h], p[:, h:]
    o_h, o_t = o[:, :h], o[:, h:]
    if combine == "spo":
        a = (s_h * p_f * o_t).sum(1)
        b = (s_t * p_b * o_h).sum(1)
    elif combine == "sp_":
        a = (s_h * p_f) @ o_t.t()
        b = (s_t * p_b) @ o_h.t()
    elif combine == "_po":
        a = (o_t * p_f) @ s_h.t()
        b = (o_h * p_b) @ s_t.t()
    else:
        return None
    return (a + b) / 2.0


def _cp(s, p, o, combine):
    """cp.py:13-30 — subject uses first half, object second half."""
    h = s.shape[1] // 2
    s_h, o_t = s[:, :h], o[:, h:]
    if combine == "spo":
        return (s_h * p * o_t).sum(1)
    if combine == "sp_":
        return (s_h * p) @ o_t.t()
    if combine == "_po":
        return (o_t * p) @ s_h.t()
    return None


def _rescal(s, p, o, combine):
    """rescal.py:14-52 — relation row viewed as row-major d x d mixing matrix."""
    n, d = p.shape[0], s.shape[-1]
    m = p.view(-1, d, d)
    if combine == "spo":
        return (torch.bmm(s.unsqueeze(1), m).view(n, d) * o).sum(-1)
    if combine == "sp_":
        return torch.bmm(s.unsqueeze(1), m).view(n, d) @ o.t()
    if combine == "_po":
        return torch.bmm(m, o.unsqueeze(2)).view(n, d) @ s.t()
    return None


def _transe(s, p, o, combine, l_norm):
    """transe.py:15-37 — spo via pairwise_distance (eps=1e-6 added to the difference),
    sp_/_po via cdist in the non-matmul form."""
    if combine == "spo":
        return -torch.nn.functional.pairwise_distance(s + p, o, p=l_norm)
    if combine == "sp_":
        return -torch.cdist(s + p, o, p=l_norm, compute_mode="donot_use_mm_for_euclid_dist")
    if combine == "_po":
        return -torch.cdist(o - p, s, p=l_norm, compute_mode="donot_use_mm_for_euclid_dist")
    return None


def _lp_nonneg(x, dim, l_norm):
    """rotate.py:205-213."""
    return x.sum(dim) if l_norm == 1.0 else torch.norm(x, dim=dim, p=l_norm)


def _rotate(s, p, o, combine, l_norm, chunk: int = 64):
    """rotate.py:20-69 + helpers :146-213.

    The reference materialises [n, E, D/2] intermediates; we do the same arithmetic
    (cos/sin, complex Hadamard, pairwise complex difference, modulus, Lp) but walk the
    query rows in chunks so the oracle fits in memory at test sizes.
    """
    h = s.shape[1] // 2
    s_re, s_im = s[:, :h], s[:, h:]
    o_re, o_im = o[:, :h], o[:, h:]
    c, sn = torch.cos(p), torch.sin(p)
    if combine == "spo":
        q_re = s_re * c - s_im * sn
        q_im = s_re * sn + s_im * c
        mod = torch.sqrt((q_re - o_re) ** 2 + (q_im - o_im) ** 2)
        return -_lp_nonneg(mod, 1, l_norm)
    if combine == "sp_":
        q_re = s_re * c - s_im * sn
        q_im = s_re * sn + s_im * c
        t_re, t_im = o_re, o_im
    elif combine == "_po":
        sn = -sn  # conjugate: inverse rotation of the tail (rotate.py:53-57)
        q_re = c * o_re - sn * o_im
        q_im = c * o_im + sn * o_re
        t_re, t_im = s_re, s_im
    else:
        return None
    out = []
    for i in range(0, q_re.shape[0], chunk):
        d_re = q_re[i : i + chunk].unsqueeze(1) - t_re
        d_im = q_im[i : i + chunk].unsqueeze(1) - t_im
        out.append(-_lp_nonneg(torch.sqrt(d_re * d_re + d_im * d_im), 2, l_norm))
    return torch.cat(out, 0) if out else q_re.new_zeros((0, t_re.shape[0]))


def score_emb(model: str, s_emb, p_emb, o_emb, combine: str, l_norm: float = 1.0):
    """RelationalScorer.score_emb for the in-scope models (kge_model.py:151-213).

    combine in {spo, sp_, _po, s_o}; output [n] for spo (viewed [n,1] by the reference,
    flattened by KgeModel.score_spo :680) else [n, m].
    """
    fn = {
        "complex": _complex,
        "distmult": _distmult,
        "simple": _simple,
        "cp": _cp,
        "rescal": _rescal,
    }.get(model)
    if fn is not None:
        out = fn(s_emb, p_emb, o_emb, combine)
    elif model == "transe":
        out = _transe(s_emb, p_emb, o_emb, combine, l_norm)
    elif model == "rotate":
        out = _rotate(s_emb, p_emb, o_emb, combine, l_norm)
    else:
        raise ValueError(f"unknown model {model}")
    if out is None:  # generic fallback, only reached for s_o by these models
        se, pe, oe, n = _generic_expand(s_emb, p_emb, o_emb, combine)
        out = score_emb(model, se, pe, oe, "spo", l_norm).view(n, -1)
    return out


# --------------------------------------------------------------------------- a3
def score_spo(model, ent, rel, s, p, o, l_norm=1.0):
    """KgeModel.score_spo kge_model.py:663-680."""
    return score_emb(model, embed(ent, s), embed(rel, p), embed(ent, o), "spo", l_norm).view(-1)


def score_sp(model, ent, rel, s, p, o=None, l_norm=1.0):
    """KgeModel.score_sp kge_model.py:682-702."""
    targets = embed_all(ent) if o is None else embed(ent, o)
    return score_emb(model, embed(ent, s), embed(rel, p), targets, "sp_", l_norm)


def score_po(model, ent, rel, p, o, s=None, l_norm=1.0):
    """KgeModel.score_po kge_model.py:704-725."""
    targets = embed_all(ent) if s is None else embed(ent, s)
    return score_emb(model, targets, embed(rel, p), embed(ent, o), "_po", l_norm)


def score_so(model, ent, rel, s, o, p=None, l_norm=1.0):
    """KgeModel.score_so kge_model.py:727-747 (generic s_o expansion for all in-scope models)."""
    targets = embed_all(rel) if p is None else embed(rel, p)
    return score_emb(model, embed(ent, s), targets, embed(ent, o), "s_o", l_norm)


def score_sp_po(model, ent, rel, s, p, o, entity_subset=None, l_norm=1.0):
    """KgeModel.score_sp_po kge_model.py:749-789: [sp | po] concatenated along columns."""
    cand = embed_all(ent) if entity_subset is None else embed(ent, entity_subset)
    pe = embed(rel, p)
    sp = score_emb(model, embed(ent, s), pe, cand, "sp_", l_norm)
    po = score_emb(model, cand, pe, embed(ent, o), "_po", l_norm)
    return torch.cat([sp, po], 1)


# --------------------------------------------------------------------------- a12 / a13
def _labels_as_matrix(scores, labels):
    """loss.py:105-117."""
    if labels.dim() == 2:
        return labels
    m = torch.zeros_like(scores)
    m[torch.arange(scores.shape[0]), labels.long()] = 1.0
    return m


def bce_loss(scores, labels, offset: float = 0.0):
    """BCEWithLogitsKgeLoss (bce_type=None): sum over all entries of
    softplus(x+off) - y*(x+off).  loss.py:153-159."""
    y = _labels_as_matrix(scores, labels).to(scores.dtype)
    x = scores + offset if offset != 0.0 else scores
    return torch.nn.functional.binary_cross_entropy_with_logits(
        x.reshape(-1), y.reshape(-1), reduction="sum"
    )


def kl_loss(scores, labels):
    """KLDivWithSoftmaxKgeLoss: index labels -> cross entropy (sum); matrix labels ->
    KLDiv(log_softmax(x), y/||y||_1) (sum).  loss.py:198-213."""
    if labels.dim() == 1:
        return torch.nn.functional.cross_entropy(scores, labels.long(), reduction="sum")
    target = torch.nn.functional.normalize(labels.to(scores.dtype), p=1, dim=1)
    return torch.nn.functional.kl_div(
        torch.log_softmax(scores, dim=1), target, reduction="sum"
    )


def kvsall_smooth_labels(labels, eps: float):
    """KvsAll label smoothing for entity targets: (1-eps)*y + 1/E. train_KvsAll.py:260-266."""
    return (1.0 - eps) * labels + 1.0 / labels.shape[1]


def train_1vsall_forward(model, ent, rel, triples, loss="bce", offset=0.0, l_norm=1.0):
    """One 1vsAll forward step: [loss(score_sp, o) + loss(score_po, s)] / batch_size with sum
    reductions.  train_1vsAll.py:48-82."""
    s, p, o = triples[:, S], triples[:, P], triples[:, O]
    fn = (lambda x, y: bce_loss(x, y, offset)) if loss == "bce" else kl_loss
    n = triples.shape[0]
    l_sp = fn(score_sp(model, ent, rel, s, p, l_norm=l_norm), o) / n
    l_po = fn(score_po(model, ent, rel, p, o, l_norm=l_norm), s) / n
    return l_sp + l_po


# --------------------------------------------------------------------------- a14 / a15
def ns_score(model, ent, rel, triples, negatives, slot, implementation="triple", l_norm=1.0):
    """BatchNegativeSample.score sampler.py:263-344 (+_score_unique_targets :347-356).

    triples [n,3], negatives [n,K] ids for `slot`; returns [n,K].
    """
    n, k = negatives.shape
    if implementation == "triple":
        t = triples.repeat(1, k).view(-1, 3).clone()
        t[:, slot] = negatives.reshape(-1)
        return score_spo(model, ent, rel, t[:, S], t[:, P], t[:, O], l_norm).view(n, k)
    if implementation == "all":
        uniq, cols = None, negatives.reshape(-1)
    else:  # batch
        uniq, cols = torch.unique(negatives.reshape(-1), return_inverse=True)
    if slot == S:
        allsc = score_po(model, ent, rel, triples[:, P], triples[:, O], uniq, l_norm)
    elif slot == P:
        allsc = score_so(model, ent, rel, triples[:, S], triples[:, O], uniq, l_norm)
    else:
        allsc = score_sp(model, ent, rel, triples[:, S], triples[:, P], uniq, l_norm)
    rows = torch.arange(n).unsqueeze(1).repeat(1, k).view(-1)
    return allsc[rows, cols].view(n, k)


def ns_scores_with_positive(model, ent, rel, triples, negatives, slot,
                            implementation="triple", l_norm=1.0):
    """TrainingJobNegativeSampling score assembly: column 0 = positive (score_spo),
    columns 1.. = negatives.  train_negative_sampling.py:139-148."""
    n, k = negatives.shape
    out = torch.empty((n, 1 + k), dtype=ent.dtype)
    out[:, 0] = score_spo(model, ent, rel, triples[:, S], triples[:, P], triples[:, O], l_norm)
    out[:, 1:] = ns_score(model, ent, rel, triples, negatives, slot, implementation, l_norm)
    return out


def ns_labels(n, k, dtype=torch.float32):
    """train_negative_sampling.py:128-137."""
    y = torch.zeros((n, 1 + k), dtype=dtype)
    y[:, 0] = 1
    return y


# --------------------------------------------------------------------------- a16 / a17
def ranks_and_ties(scores, true_scores, rtol=1e-4, atol=1e-5) -> Tuple[torch.Tensor, torch.Tensor]:
    """EntityRankingJob._get_ranks_and_num_ties eval_entity_ranking.py:571-596.

    NaN -> -inf on both sides; close = |x - t| <= atol + rtol*|t| (torch.isclose, which
    also treats equal infinities as close); rank = #(x > t and not close); ties = #close.
    """
    x = scores.clone()
    x[torch.isnan(x)] = float("-inf")
    t = true_scores.clone().view(-1)
    t[torch.isnan(t)] = float("-inf")
    close = torch.isclose(x, t.view(-1, 1), rtol=rtol, atol=atol)
    greater = x > t.view(-1, 1)
    return (greater & ~close).sum(1, dtype=torch.long), close.sum(1, dtype=torch.long)


def filter_and_rank(scores_sp, scores_po, labels, o_true, s_true, rtol=1e-4, atol=1e-5):
    """EntityRankingJob._filter_and_rank eval_entity_ranking.py:533-569.

    labels: dense [n, 2*chunk] with +inf at known-true columns (own answer zeroed), or None.
    Returns s_rank, s_ties, o_rank, o_ties.
    """
    c = scores_sp.shape[1]
    if labels is not None:
        scores_sp = scores_sp - labels[:, :c]
        scores_po = scores_po - labels[:, c:]
    o_rank, o_ties = ranks_and_ties(scores_sp, o_true, rtol, atol)
    s_rank, s_ties = ranks_and_ties(scores_po, s_true, rtol, atol)
    return s_rank, s_ties, o_rank, o_ties


def final_ranks(rank, ties, tie_handling="rounded_mean_rank"):
    """EntityRankingJob._get_ranks eval_entity_ranking.py:598-618."""
    if tie_handling == "rounded_mean_rank":
        return rank + ties // 2
    if tie_handling == "best_rank":
        return rank
    if tie_handling == "worst_rank":
        return rank + ties - 1
    raise NotImplementedError(tie_handling)


def true_scores_sp_path(model, ent, rel, s, p, o, l_norm=1.0):
    """True-triple scores computed with the sp_/_po code path (not spo) so tie handling is
    consistent: score_sp(s,p,unique_o) then gather. eval_entity_ranking.py:192-203."""
    uo, oi = torch.unique(o, return_inverse=True)
    o_true = score_sp(model, ent, rel, s, p, uo, l_norm)[torch.arange(len(s)), oi]
    us, si = torch.unique(s, return_inverse=True)
    s_true = score_po(model, ent, rel, p, o, us, l_norm)[torch.arange(len(s)), si]
    return o_true, s_true


def entity_ranking_metrics(model, ent, rel, eval_triples, filter_splits, test_triples=None, l_norm=1.0,
                           tie_handling="rounded_mean_rank", hits_at_k=(1, 3, 10), rtol=1e-4, atol=1e-5):
    """EntityRankingJob._evaluate restated at small scale (eval_entity_ranking.py:103-487): for every
    evaluation triple rank the true object among all objects (sp_) and the true subject among all
    subjects (_po); raw, filtered (known-true answers from `filter_splits` get +inf subtracted, the query's
    own answer excepted :169-182,287-290) and, if `test_triples` is given, filtered_with_test (the test
    split's answers filtered ON TOP of the already filtered scores :278-303).  Ranks are 0-based, metrics come
    from the rank histogram over both directions (:620-649).  Returns {metric: value} with the reference's
    keys."""
    E = ent.shape[0]
    s, p, o = eval_triples[:, S].long(), eval_triples[:, P].long(), eval_triples[:, O].long()
    n = len(s)
    scores = score_sp_po(model, ent, rel, s, p, o, None, l_norm)
    sp, po = scores[:, :E], scores[:, E:]
    o_true, s_true = sp[torch.arange(n), o], po[torch.arange(n), s]

    def label_matrix(known):
        known = known.long()
        labels = torch.zeros((n, 2 * E), dtype=scores.dtype)
        for i in range(n):
            m_sp = (known[:, S] == s[i]) & (known[:, P] == p[i])
            labels[i, known[m_sp, O]] = float("inf")
            m_po = (known[:, P] == p[i]) & (known[:, O] == o[i])
            labels[i, E + known[m_po, S]] = float("inf")
        labels[torch.arange(n), o] = 0.0
        labels[torch.arange(n), E + s] = 0.0
        return labels

    rankings = [("", None), ("_filtered", label_matrix(torch.cat([t.long() for t in filter_splits], 0)))]
    if test_triples is not None:
        rankings.append(("_filtered_with_test", label_matrix(test_triples)))
    out = {}
    for suffix, lab in rankings:
        s_rank, s_ties, o_rank, o_ties = filter_and_rank(sp, po, lab, o_true, s_true, rtol, atol)
        if lab is not None:          # "from now on, use filtered scores" :298-300
            sp, po = sp - lab[:, :E], po - lab[:, E:]
        ranks = torch.cat([final_ranks(s_rank, s_ties, tie_handling), final_ranks(o_rank, o_ties, tie_handling)])
        r1 = (ranks + 1).double()
        out["mean_rank" + suffix] = float(r1.mean())
        out["mean_reciprocal_rank" + suffix] = float((1.0 / r1).mean())
        for k in hits_at_k:
            out[f"hits_at_{k}{suffix}"] = float((ranks < k).double().mean())
    return out


# --------------------------------------------------------------------------- f-4: reciprocal relations
def reciprocal_score_spo(model, ent, rel2, s, p, o, direction, num_relations, l_norm=1.0):
    """ReciprocalRelationsModel.score_spo (reciprocal_relations_model.py:72-82); rel2 has 2R rows."""
    if direction == "o":
        return score_spo(model, ent, rel2, s, p, o, l_norm)
    if direction == "s":
        return score_spo(model, ent, rel2, o, p + num_relations, s, l_norm)
    raise Exception("The reciprocal relations model cannot compute undirected spo scores.")


def reciprocal_score_po(model, ent, rel2, p, o, num_relations, s_subset=None, l_norm=1.0):
    """(?, p, o) scored as the object query (o, p + R, ?)   reciprocal_relations_model.py:84-91."""
    return score_sp(model, ent, rel2, o, p + num_relations, s_subset, l_norm=l_norm)


def reciprocal_score_sp_po(model, ent, rel2, s, p, o, num_relations, entity_subset=None, l_norm=1.0):
    """reciprocal_relations_model.py:97-124."""
    return torch.cat([score_sp(model, ent, rel2, s, p, entity_subset, l_norm=l_norm),
                      reciprocal_score_po(model, ent, rel2, p, o, num_relations, entity_subset, l_norm)], 1)


# --------------------------------------------------------------------------- f-3: penalties, normalisation
def _abs_complex(w):
    """lookup_embedder.py:118-121 (modulus of the complex halves, +1e-14 under the root)."""
    h = w.shape[1] // 2
    return torch.sqrt(w[:, :h] ** 2 + w[:, h:] ** 2 + 1e-14)


def lookup_penalty(weight, regularize="lp", regularize_weight=0.0, p=2, weighted=False, indexes=None,
                   space="euclidean"):
    """LookupEmbedder.penalty (lookup_embedder.py:123-177) as a scalar tensor: Lp / N3 regularisation over the
    whole table (unweighted) or over the batch's unique rows weighted by their counts and divided by the number
    of indexes (weighted)."""
    if regularize == "" or regularize_weight == 0.0:
        return weight.new_zeros(())
    if regularize == "n3":
        p = 3
    elif regularize != "lp":
        raise ValueError(f"Invalid value regularize={regularize}")
    if not weighted:
        params = weight
        if regularize == "n3" and space == "complex":
            params = _abs_complex(params)
        return (regularize_weight / p * params.norm(p=p) ** p).sum()
    uniq, counts = torch.unique(indexes, return_counts=True)
    params = weight[uniq.long()]
    if regularize == "n3" and space == "complex":
        params = _abs_complex(params)
    if (p % 2 == 1) and regularize != "n3":
        params = torch.abs(params)
    # len(indexes): the entity call passes an [n, 2] block of subjects and objects, so the divisor is n, not 2n
    return (regularize_weight / p * (params ** p * counts.float().view(-1, 1))).sum() / indexes.shape[0]


def model_penalty(ent, rel, triples, ent_opts, rel_opts):
    """KgeModel.penalty with a shared entity embedder (kge_model.py:603-649): relation penalty on triples[:,P];
    entity penalty on the [n,2] block of subjects and objects if weighted, else the table penalty doubled."""
    total = lookup_penalty(rel, indexes=triples[:, P], **rel_opts)
    if ent_opts.get("weighted", False):
        total = total + lookup_penalty(ent, indexes=triples[:, [S, O]], **ent_opts)
    else:
        total = total + 2.0 * lookup_penalty(ent, **ent_opts)
    return total


def normalize_embeddings(weight, p):
    """LookupEmbedder._normalize_embeddings (lookup_embedder.py:64-69): rows scaled to unit Lp norm (p > 0)."""
    return torch.nn.functional.normalize(weight, p=p, dim=-1) if p > 0 else weight


# --------------------------------------------------------------------------- synthetic inputs
def make_tables(model: str, E: int, R: int, D: int, sigma: float = 1.0, seed: int = 1234,
                dtype=torch.float32):
    """Seeded synthetic embedding tables: entity/relation ~ N(0, sigma);
    RotatE relation phases ~ U(-pi, pi) (rotate.yaml:22-26)."""
    g = torch.Generator().manual_seed(seed)
    ent = torch.randn((E, D), generator=g, dtype=torch.float32) * sigma
    dr = relation_dim(model, D)
    if model == "rotate":
        rel = (torch.rand((R, dr), generator=g, dtype=torch.float32) * 2.0 - 1.0) * math.pi
    else:
        rel = torch.randn((R, dr), generator=g, dtype=torch.float32) * sigma
    return ent.to(dtype), rel.to(dtype)


def make_triples(E: int, R: int, n: int, seed: int = 0):
    g = torch.Generator().manual_seed(seed)
    return torch.stack(
        [
            torch.randint(0, E, (n,), generator=g),
            torch.randint(0, R, (n,), generator=g),
            torch.randint(0, E, (n,), generator=g),
        ],
        1,
    )
