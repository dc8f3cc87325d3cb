"""CPU restatement of the ALGEBRA the CUDA path uses (TEST INFRASTRUCTURE — not product code).

`oracle/kge_oracle.py` restates the reference's own op sequence.  This module restates, in plain torch on the
CPU, the factorisation the kernels are built on (kge_b200/csrc/fold.cuh): every 1-vs-N score is
`pair(Q_i, cand_j[cols])` with `Q = fold(a, p)` depending on the per-row operands only — and, on top of it,
the analytic backward of the fused score+loss step that the gradient kernels will implement:

    g  = dL/dz                    [nq, E]   (sigmoid(z+off) - y) / n   |  (y_sum * softmax(z) - y) / n
    dQ = g  @ cand[:, cols]       [nq, K]
    dT = g^T @ Q                  [E,  K]   (added into the entity-table gradient at `cols`)
    (da, dp) = unfold(a, p, dQ)             hand-derived vector-Jacobian products of `fold`
(for TransE / RotatE the two GEMMs become sums over the sign / direction fields of the pairwise differences,
`pair_backward`).

tests/test_fold_algebra.py checks fold+pair against the oracle's scores for all seven models, `unfold`
against autograd, and the assembled table gradients against autograd of the oracle's training step and
against gradients recorded from the live reference (tests/golden/grads_*.npz).
"""
from __future__ import annotations

import torch

S, P, O = 0, 1, 2


def cand_cols(model: str, combine: str, D: int):
    """(column offset, K) of the candidate columns the folded query is paired with (fold.cuh folded_problem)."""
    if model == "cp":
        return (D // 2, D // 2) if combine == "sp_" else (0, D // 2)
    return 0, D


def fold(model: str, combine: str, a: torch.Tensor, p: torch.Tensor) -> torch.Tensor:
    """Folded queries [n, K]; a = subject rows for sp_, object rows for _po (fold.cuh fold_element)."""
    sp = combine == "sp_"
    h = a.shape[1] // 2
    if model == "complex":
        a_re, a_im, p_re, p_im = a[:, :h], a[:, h:], p[:, :h], p[:, h:]
        if sp:
            return torch.cat([a_re * p_re - a_im * p_im, a_im * p_re + a_re * p_im], 1)
        return torch.cat([p_re * a_re + p_im * a_im, p_re * a_im - p_im * a_re], 1)
    if model == "distmult":
        return a * p
    if model == "simple":
        a_h, a_t, p_f, p_b = a[:, :h], a[:, h:], p[:, :h], p[:, h:]
        if sp:
            return 0.5 * torch.cat([a_t * p_b, a_h * p_f], 1)
        return 0.5 * torch.cat([a_t * p_f, a_h * p_b], 1)
    if model == "cp":
        return a[:, :h] * p if sp else a[:, h:] * p
    if model == "rescal":
        d = a.shape[1]
        m = p.view(-1, d, d)
        if sp:
            return torch.bmm(a.unsqueeze(1), m).squeeze(1)           # q_j = sum_i s_i M[i,j]
        return torch.bmm(m, a.unsqueeze(2)).squeeze(2)               # q_i = sum_j M[i,j] o_j
    if model == "transe":
        return a + p if sp else a - p
    if model == "rotate":
        c, sn = torch.cos(p), torch.sin(p)
        a_re, a_im = a[:, :h], a[:, h:]
        if sp:
            return torch.cat([a_re * c - a_im * sn, a_re * sn + a_im * c], 1)
        return torch.cat([c * a_re + sn * a_im, c * a_im - sn * a_re], 1)
    raise ValueError(model)


def pair_scores(model: str, Q: torch.Tensor, T: torch.Tensor, l_norm: float = 1.0) -> torch.Tensor:
    """pair(Q_i, T_j) for all i, j: dot product, -Lp distance (TransE) or -sum of complex moduli (RotatE)."""
    if model in ("complex", "distmult", "simple", "cp", "rescal"):
        return Q @ T.t()
    if model == "transe":
        d = (Q.unsqueeze(1) - T.unsqueeze(0)).abs()
        return -(d.sum(2) if l_norm == 1.0 else d.pow(l_norm).sum(2).pow(1.0 / l_norm))
    if model == "rotate":
        h = Q.shape[1] // 2
        dre = Q[:, :h].unsqueeze(1) - T[:, :h].unsqueeze(0)
        dim = Q[:, h:].unsqueeze(1) - T[:, h:].unsqueeze(0)
        mod = torch.sqrt(dre * dre + dim * dim)
        return -(mod.sum(2) if l_norm == 1.0 else mod.pow(l_norm).sum(2).pow(1.0 / l_norm))
    raise ValueError(model)


def score_1vsN(model, combine, ent, rel, q_idx, p_idx, l_norm=1.0):
    """Scores of (q, p) against all entities through the folded form."""
    off, K = cand_cols(model, combine, ent.shape[1])
    Q = fold(model, combine, ent[q_idx.long()], rel[p_idx.long()])
    return pair_scores(model, Q, ent[:, off:off + K], l_norm)


# --------------------------------------------------------------------------------------------- backward
def unfold(model: str, combine: str, a: torch.Tensor, p: torch.Tensor, dQ: torch.Tensor):
    """(da, dp): vector-Jacobian products of `fold` at (a, p) with upstream dQ — the row-wise formulas the
    gradient prologue kernels implement (dot family)."""
    sp = combine == "sp_"
    h = a.shape[1] // 2
    if model == "complex":
        a_re, a_im, p_re, p_im = a[:, :h], a[:, h:], p[:, :h], p[:, h:]
        g_re, g_im = dQ[:, :h], dQ[:, h:]
        if sp:    # Q_re = a_re p_re - a_im p_im ; Q_im = a_im p_re + a_re p_im
            da = torch.cat([g_re * p_re + g_im * p_im, -g_re * p_im + g_im * p_re], 1)
            dp = torch.cat([g_re * a_re + g_im * a_im, -g_re * a_im + g_im * a_re], 1)
        else:     # Q_re = p_re a_re + p_im a_im ; Q_im = p_re a_im - p_im a_re
            da = torch.cat([g_re * p_re - g_im * p_im, g_re * p_im + g_im * p_re], 1)
            dp = torch.cat([g_re * a_re + g_im * a_im, g_re * a_im - g_im * a_re], 1)
        return da, dp
    if model == "distmult":
        return dQ * p, dQ * a
    if model == "simple":
        a_h, a_t, p_f, p_b = a[:, :h], a[:, h:], p[:, :h], p[:, h:]
        g0, g1 = 0.5 * dQ[:, :h], 0.5 * dQ[:, h:]
        if sp:    # Q = 1/2 [a_t p_b | a_h p_f]
            return torch.cat([g1 * p_f, g0 * p_b], 1), torch.cat([g1 * a_h, g0 * a_t], 1)
        # Q = 1/2 [a_t p_f | a_h p_b]
        return torch.cat([g1 * p_b, g0 * p_f], 1), torch.cat([g0 * a_t, g1 * a_h], 1)
    if model == "cp":
        z = torch.zeros_like(a[:, :h])
        if sp:    # Q = a[:h] p
            return torch.cat([dQ * p, z], 1), dQ * a[:, :h]
        return torch.cat([z, dQ * p], 1), dQ * a[:, h:]
    if model == "rescal":
        d = a.shape[1]
        m = p.view(-1, d, d)
        if sp:    # q = a^T M : da = M dq, dM = a dq^T
            da = torch.bmm(m, dQ.unsqueeze(2)).squeeze(2)
            dm = torch.bmm(a.unsqueeze(2), dQ.unsqueeze(1))
        else:     # q = M a : da = M^T dq, dM = dq a^T
            da = torch.bmm(dQ.unsqueeze(1), m).squeeze(1)
            dm = torch.bmm(dQ.unsqueeze(2), a.unsqueeze(1))
        return da, dm.reshape(-1, d * d)
    if model == "transe":      # Q = a + p  |  Q = a - p
        return dQ, (dQ if sp else -dQ)
    if model == "rotate":      # Q = a * e^{i theta}  |  Q = conj(e^{i theta}) * a ;  p = theta [n, h]
        c, sn = torch.cos(p), torch.sin(p)
        a_re, a_im = a[:, :h], a[:, h:]
        g_re, g_im = dQ[:, :h], dQ[:, h:]
        if sp:    # Q_re = a_re c - a_im s ; Q_im = a_re s + a_im c
            da = torch.cat([g_re * c + g_im * sn, -g_re * sn + g_im * c], 1)
            dp = g_re * (-a_re * sn - a_im * c) + g_im * (a_re * c - a_im * sn)
        else:     # Q_re = c a_re + s a_im ; Q_im = c a_im - s a_re
            da = torch.cat([g_re * c - g_im * sn, g_re * sn + g_im * c], 1)
            dp = g_re * (-sn * a_re + c * a_im) + g_im * (-sn * a_im - c * a_re)
        return da, dp
    raise ValueError(model)


def pair_backward(model: str, Q: torch.Tensor, T: torch.Tensor, g: torch.Tensor, l_norm: float = 1.0):
    """(dQ, dT) of sum_ij g_ij * pair(Q_i, T_j): two GEMMs for the dot family; for the distance family the
    sign / direction fields of the pairwise differences (what a CUDA-core gradient kernel accumulates tile by
    tile, never materialising [n, E, D])."""
    if model in ("complex", "distmult", "simple", "cp", "rescal"):
        return g @ T, g.t() @ Q
    if model == "transe":      # z = -||q - t||_p
        d = Q.unsqueeze(1) - T.unsqueeze(0)                       # [n, E, K]
        if l_norm == 1.0:
            w = -torch.sign(d)
        else:
            nrm = d.abs().pow(l_norm).sum(2, keepdim=True).pow(1.0 / l_norm)
            w = -torch.sign(d) * d.abs().pow(l_norm - 1) / nrm.clamp_min(1e-30).pow(l_norm - 1)
        gw = g.unsqueeze(2) * w
        return gw.sum(1), -gw.sum(0)
    if model == "rotate":      # z = -sum_k |q_k - t_k| (complex modulus), L1 over k
        if l_norm != 1.0:
            raise NotImplementedError("rotate backward is restated for l_norm = 1 only")
        h = Q.shape[1] // 2
        dre = Q[:, :h].unsqueeze(1) - T[:, :h].unsqueeze(0)
        dim = Q[:, h:].unsqueeze(1) - T[:, h:].unsqueeze(0)
        mod = torch.sqrt(dre * dre + dim * dim).clamp_min(1e-30)
        wre, wim = -dre / mod, -dim / mod
        gre, gim = g.unsqueeze(2) * wre, g.unsqueeze(2) * wim
        return torch.cat([gre.sum(1), gim.sum(1)], 1), -torch.cat([gre.sum(0), gim.sum(0)], 1)
    raise ValueError(model)


def loss_grad(z: torch.Tensor, labels: torch.Tensor, loss: str, offset: float, batch_size: int) -> torch.Tensor:
    """dL/dz of KgeLoss(z, labels) / batch_size (loss.py:105-117,150-157 BCE with offset; :198-213 KL);
    labels: positions [n] or matrix [n, E]."""
    if labels.dim() == 1:
        y = torch.zeros_like(z)
        y[torch.arange(z.shape[0]), labels.long()] = 1.0
    else:
        y = labels.to(z.dtype)
    if loss == "bce":
        return (torch.sigmoid(z + offset) - y) / batch_size
    if loss == "kl":
        # the reference normalises the labels first: KLDiv(log_softmax(z), y / max(||y||_1, 1e-12))  loss.py:209-213
        yn = y / y.sum(1, keepdim=True).clamp_min(1e-12)
        return (yn.sum(1, keepdim=True) * torch.softmax(z, 1) - yn) / batch_size
    raise ValueError(loss)


def train_1vsall_backward(model, ent, rel, triples, loss="bce", offset=0.0, l_norm=1.0):
    """(dEnt [E,D], dRel [R,Dr]) of [loss(score_sp, o) + loss(score_po, s)] / n (train_1vsAll.py:48-82),
    assembled as the gradient kernels will: G pass, pair backward (two GEMMs for the dot family), row-wise
    unfold, scatter-add."""
    s, p, o = triples[:, S].long(), triples[:, P].long(), triples[:, O].long()
    n, D = triples.shape[0], ent.shape[1]
    d_ent, d_rel = torch.zeros_like(ent), torch.zeros_like(rel)
    for combine, q_idx, lab in (("sp_", s, o), ("_po", o, s)):
        off, K = cand_cols(model, combine, D)
        a, pr = ent[q_idx], rel[p]
        Q = fold(model, combine, a, pr)
        T = ent[:, off:off + K]
        g = loss_grad(pair_scores(model, Q, T, l_norm), lab, loss, offset, n)
        dQ, dT = pair_backward(model, Q, T, g, l_norm)
        d_ent[:, off:off + K] += dT
        da, dp = unfold(model, combine, a, pr, dQ)
        d_ent.index_add_(0, q_idx, da)
        d_rel.index_add_(0, p, dp)
    return d_ent, d_rel


def spo_backward(model, ent, rel, s, p, o, g, d_ent, d_rel, l_norm=1.0):
    """Accumulate into (d_ent, d_rel) the gradient of sum_i g_i * score(s_i, p_i, o_i) through the folded form
    score = pair(fold_sp(s, p), o[cols]): the row-wise backward that covers positives and every negative-sample
    slot (a negative sample is one more triple with one slot replaced)."""
    s, p, o = s.long(), p.long(), o.long()
    off, K = cand_cols(model, "sp_", ent.shape[1])
    a, pr = ent[s], rel[p]
    Q = fold(model, "sp_", a, pr)
    T = ent[o][:, off:off + K]
    if model in ("complex", "distmult", "simple", "cp", "rescal"):
        dQ, dT = g.unsqueeze(1) * T, g.unsqueeze(1) * Q
    elif model == "transe":
        d = Q - T
        if l_norm == 1.0:
            w = -torch.sign(d)
        else:
            nrm = d.abs().pow(l_norm).sum(1, keepdim=True).pow(1.0 / l_norm)
            w = -torch.sign(d) * d.abs().pow(l_norm - 1) / nrm.clamp_min(1e-30).pow(l_norm - 1)
        dQ = g.unsqueeze(1) * w
        dT = -dQ
    elif model == "rotate":
        h = Q.shape[1] // 2
        dre, dim = Q[:, :h] - T[:, :h], Q[:, h:] - T[:, h:]
        mod = torch.sqrt(dre * dre + dim * dim).clamp_min(1e-30)
        dQ = g.unsqueeze(1) * torch.cat([-dre / mod, -dim / mod], 1)
        dT = -dQ
    else:
        raise ValueError(model)
    da, dp = unfold(model, "sp_", a, pr, dQ)
    d_ent.index_add_(0, s, da)
    d_rel.index_add_(0, p, dp)
    full = torch.zeros((o.numel(), ent.shape[1]), dtype=ent.dtype)
    full[:, off:off + K] = dT
    d_ent.index_add_(0, o, full)


def ns_backward(model, ent, rel, triples, negatives, offset=0.0, l_norm=1.0):
    """(dEnt, dRel) of one negative-sampling batch with BCE (train_negative_sampling.py:113-164): per slot the
    [n, 1+K] block (positive first, label 1; negatives label 0), loss summed and divided by the batch size.
    negatives = {slot: [n, K] ids}."""
    n = triples.shape[0]
    d_ent, d_rel = torch.zeros_like(ent), torch.zeros_like(rel)
    for slot, neg in negatives.items():
        k = neg.shape[1]
        t = triples.long().repeat_interleave(1 + k, 0).view(n, 1 + k, 3).clone()
        t[:, 1:, slot] = neg.long()
        t = t.view(-1, 3)
        z = pair_rowwise(model, ent, rel, t[:, S], t[:, P], t[:, O], l_norm)
        y = torch.zeros((n, 1 + k), dtype=ent.dtype)
        y[:, 0] = 1.0
        g = ((torch.sigmoid(z + offset) - y.view(-1)) / n)
        spo_backward(model, ent, rel, t[:, S], t[:, P], t[:, O], g, d_ent, d_rel, l_norm)
    return d_ent, d_rel


def pair_rowwise(model, ent, rel, s, p, o, l_norm=1.0):
    """score(s_i, p_i, o_i) through the folded form (row-wise pair)."""
    off, K = cand_cols(model, "sp_", ent.shape[1])
    Q = fold(model, "sp_", ent[s.long()], rel[p.long()])
    T = ent[o.long()][:, off:off + K]
    if model in ("complex", "distmult", "simple", "cp", "rescal"):
        return (Q * T).sum(1)
    if model == "transe":
        d = (Q - T).abs()
        return -(d.sum(1) if l_norm == 1.0 else d.pow(l_norm).sum(1).pow(1.0 / l_norm))
    h = Q.shape[1] // 2
    mod = torch.sqrt((Q[:, :h] - T[:, :h]) ** 2 + (Q[:, h:] - T[:, h:]) ** 2)
    return -(mod.sum(1) if l_norm == 1.0 else mod.pow(l_norm).sum(1).pow(1.0 / l_norm))
