"""TEST INFRASTRUCTURE: an oracle-backed stand-in for kge_b200.engine so that the HOST logic of the plugin
(kge_b200/plugin: model overrides, autograd wrappers, job subclasses — CSR construction, sub-batch scaling,
fall-backs) can be exercised on a machine without a GPU.  The CUDA path itself is covered by the `-m gpu` tests
(tests/test_gpu_jobs.py runs the same jobs on the real engine)."""
from __future__ import annotations

import contextlib

import torch

from oracle import kge_oracle as orc


def _rows(tab, idx):
    return tab if idx is None else tab[idx.long()]


def score_spo(model, ent_s, rel, ent_o, s=None, p=None, o=None, l_norm=1.0):
    return orc.score_emb(model, _rows(ent_s, s), _rows(rel, p), _rows(ent_o, o), "spo", l_norm).view(-1)


def score_1vsN(model, combine, q_tab, rel, cand_tab, q=None, p=None, cand=None, l_norm=1.0, precision="auto", out=None):
    a, r, c = _rows(q_tab, q), _rows(rel, p), _rows(cand_tab, cand)
    if combine == "sp_":
        return orc.score_emb(model, a, r, c, "sp_", l_norm)
    return orc.score_emb(model, c, r, a, "_po", l_norm)


def score_sp_po(model, ent, rel, s, p, o, entity_subset=None, l_norm=1.0, precision="auto"):
    return orc.score_sp_po(model, ent, rel, s.long(), p.long(), o.long(), entity_subset, l_norm)


def train_1vsall_forward(model, ent, rel, triples, loss="bce", offset=0.0, l_norm=1.0, precision="auto", out=None,
                         workspace=None):
    return orc.train_1vsall_forward(model, ent, rel, triples.long(), loss, offset, l_norm)


def score_1vsN_loss(model, combine, q_tab, rel, cand_tab, labels, q=None, p=None, cand=None, loss="bce", offset=0.0,
                    l_norm=1.0, precision="auto", return_rows=False):
    x = score_1vsN(model, combine, q_tab, rel, cand_tab, q, p, cand, l_norm)
    return orc.bce_loss(x, labels, offset) if loss == "bce" else orc.kl_loss(x, labels)


def score_1vsN_loss_csr(model, combine, q_tab, rel, cand_tab, csr_offsets, csr_cols, q=None, p=None, loss="kl",
                        offset=0.0, label_smoothing=0.0, l_norm=1.0, precision="auto", return_rows=False):
    x = score_1vsN(model, combine, q_tab, rel, cand_tab, q, p, None, l_norm)
    n, m = x.shape
    y = torch.zeros((n, m))
    counts = csr_offsets[1:] - csr_offsets[:-1]
    rows = torch.repeat_interleave(torch.arange(n), counts)
    y.index_put_((rows, csr_cols.long()), torch.ones(len(rows)), accumulate=True)   # duplicates add up
    if label_smoothing > 0:
        y = orc.kvsall_smooth_labels(y, label_smoothing)
    return orc.bce_loss(x, y, offset) if loss == "bce" else orc.kl_loss(x, y)


def ns_score(model, ent, rel, triples, negatives, slot, with_positive=False, l_norm=1.0):
    if with_positive:
        return orc.ns_scores_with_positive(model, ent, rel, triples.long(), negatives.long(), slot, l_norm=l_norm)
    return orc.ns_score(model, ent, rel, triples.long(), negatives.long(), slot, l_norm=l_norm)


def loss_dense(scores, labels, loss="bce", offset=0.0, return_rows=False):
    return orc.bce_loss(scores, labels, offset) if loss == "bce" else orc.kl_loss(scores, labels)


def train_1vsall_backward(model, ent, rel, triples, loss="bce", offset=0.0, l_norm=1.0):
    from oracle import kge_fold as kf

    return kf.train_1vsall_backward(model, ent.detach(), rel.detach(), triples.long(), loss, offset, l_norm)


def ns_backward(model, ent, rel, triples, negatives, offset=0.0, l_norm=1.0, batch_size=None):
    from oracle import kge_fold as kf

    d_ent, d_rel = kf.ns_backward(model, ent.detach(), rel.detach(), triples.long(), negatives, offset, l_norm)
    scale = triples.shape[0] / float(batch_size or triples.shape[0])     # the oracle divides by n, the job by the batch
    return d_ent * scale, d_rel * scale


def score_1vsN_backward(model, combine, ent, rel, q, p, grad_scores, l_norm=1.0):
    e, r = ent.detach().clone().requires_grad_(True), rel.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        x = score_1vsN(model, combine, e, r, e, q, p, None, l_norm)
        return torch.autograd.grad(x, (e, r), grad_scores)


def score_1vsN_loss_csr_backward(model, combine, ent, rel, q, p, csr_offsets, csr_cols, loss="kl", offset=0.0,
                                 label_smoothing=0.0, batch_size=None):
    e, r = ent.detach().clone().requires_grad_(True), rel.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        val = score_1vsN_loss_csr(model, combine, e, r, e, csr_offsets, csr_cols, q, p, loss, offset, label_smoothing)
        return torch.autograd.grad(val / (batch_size or q.numel()), (e, r))


class Step1vsAll:
    """Stand-in of engine.Step1vsAll (prepared fused step)."""

    def __init__(self, model, ent, rel, max_n, loss="bce", offset=0.0, l_norm=1.0, precision="auto"):
        self.model, self.ent, self.rel, self.max_n = model, ent, rel, max_n
        self.loss, self.offset, self.l_norm = loss, offset, l_norm

    def matches(self, ent, rel, n):
        return n <= self.max_n and ent.data_ptr() == self.ent.data_ptr() and rel.data_ptr() == self.rel.data_ptr()

    def __call__(self, triples):
        _counter["n"] += 1
        return orc.train_1vsall_forward(self.model, self.ent, self.rel, triples.long(), self.loss, self.offset, self.l_norm)

    def call_host(self, triples_host):
        return float(self(triples_host))


_counter = {"n": 0}


def launch_count(reset=False):
    v = _counter["n"]
    if reset:
        _counter["n"] = 0
    return v


@contextlib.contextmanager
def installed():
    """Swap the functions of kge_b200.engine for the oracle-backed ones above for the duration of the block."""
    from kge_b200 import engine

    names = ["score_spo", "score_1vsN", "score_sp_po", "train_1vsall_forward", "score_1vsN_loss",
             "score_1vsN_loss_csr", "ns_score", "loss_dense", "train_1vsall_backward", "ns_backward", "score_1vsN_backward", "score_1vsN_loss_csr_backward", "launch_count"]
    saved = {k: getattr(engine, k) for k in names}
    g = globals()

    def counted(fn):
        def wrapper(*a, **kw):
            _counter["n"] += 1
            return fn(*a, **kw)
        return wrapper

    for k in names:
        setattr(engine, k, g[k] if k == "launch_count" else counted(g[k]))
    saved["Step1vsAll"] = engine.Step1vsAll
    engine.Step1vsAll = Step1vsAll
    try:
        yield
    finally:
        for k, v in saved.items():
            setattr(engine, k, v)
