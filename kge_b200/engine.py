"""Tensor-level façade over the C ABI: torch CUDA tensors in, torch CUDA tensors out.

torch is used for device memory (caching allocator), streams and nothing else: every number is
produced by libb200kge's hand-written sm_100a kernels.  All functions raise on CPU tensors.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import LOSS, MODELS, PREC, Labels, Rows, SP_, _PO

S, P, O = 0, 1, 2


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "kge_b200 runs on CUDA (sm_100) tensors only; there is no CPU path "
                f"(got a tensor on {t.device})"
            )


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"expected float32 embeddings, got {t.dtype}")
    if t.dim() != 2:
        raise ValueError(f"expected a 2-D embedding matrix, got shape {tuple(t.shape)}")
    if t.stride(1) != 1:
        t = t.contiguous()
    return t


def _i64(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if t.dim() != 1:
        t = t.reshape(-1)
    if t.dtype != torch.int64 or not t.is_contiguous():
        t = t.long().contiguous()  # lookup_embedder.py:97 does indexes.long()
    return t


class _Keep:
    """Keeps tensors alive for the duration of a call (the C side borrows raw pointers)."""

    def __init__(self):
        self.refs = []

    def rows(self, base: torch.Tensor, idx: Optional[torch.Tensor] = None) -> Rows:
        base = _f32(base)
        idx = _i64(idx)
        self.refs += [base, idx]
        r = Rows()
        r.base = base.data_ptr()
        r.idx = idx.data_ptr() if idx is not None else None
        r.rows = idx.numel() if idx is not None else base.shape[0]
        r.ld = base.stride(0) if base.shape[0] > 1 else max(base.shape[1], base.stride(0))
        r.dim = base.shape[1]
        return r


def _stream(dev) -> C.c_void_p:
    """The operands' current stream.  The library launches on the CURRENT device, so follow the operands when they live
    elsewhere (`job.device: cuda:1` in a process that never called set_device, e.g. LibKGE's search workers,
    kge/job/search.py:36-40); a no-op when the devices already agree."""
    idx = dev.index if isinstance(dev, torch.device) else torch.device(dev).index
    if idx is not None and idx != torch.cuda.current_device():
        torch.cuda.set_device(idx)
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _workspace(model_id: int, n: int, m: int, D: int, has_idx: bool, dev) -> torch.Tensor:
    nbytes = _lib.load().b200kge_workspace_bytes(model_id, n, m, D, 1 if has_idx else 0)
    # torch's caching allocator raises torch.cuda.OutOfMemoryError ("CUDA out of memory") on failure
    return torch.empty(nbytes, dtype=torch.uint8, device=dev)


def device_ok() -> bool:
    return _lib.load().b200kge_device_ok() == 0


def launch_count(reset: bool = False) -> int:
    return int(_lib.load().b200kge_launch_count(1 if reset else 0))


def profile_enable(on: bool = True) -> None:
    _lib.load().b200kge_profile_enable(1 if on else 0)


def profile_last_ms() -> float:
    ms = C.c_float(0.0)
    _lib.check(_lib.load().b200kge_profile_last_ms(C.byref(ms)))
    return float(ms.value)


# ------------------------------------------------------------------------------------------------
def score_spo(model: str, ent_s, rel, ent_o, s=None, p=None, o=None, l_norm: float = 1.0):
    """Row-wise scores.  With indexes: tables + gather fused (KgeModel.score_spo); without:
    already-gathered embeddings (RelationalScorer.score_emb(..., "spo"))."""
    _require_cuda(ent_s, rel, ent_o)
    lib, k = _lib.load(), _Keep()
    rs, rp, ro = k.rows(ent_s, s), k.rows(rel, p), k.rows(ent_o, o)
    n = int(rs.rows)
    if rp.rows != n or ro.rows != n:
        raise ValueError("spo scoring needs the same number of s, p and o rows")
    out = torch.empty(n, dtype=torch.float32, device=ent_s.device)
    _lib.check(lib.b200kge_score_spo(MODELS[model], l_norm, C.byref(rs), C.byref(rp), C.byref(ro), n,
                                     out.data_ptr(), _stream(ent_s.device)))
    return out


def score_1vsN(model: str, combine: str, q_tab, rel, cand_tab, q=None, p=None, cand=None,
               l_norm: float = 1.0, precision: str = "auto", out: Optional[torch.Tensor] = None):
    """[n, m] scores of n (entity, relation) rows against m candidate entities.

    combine "sp_": q rows are subjects, candidates are objects; "_po": q rows are objects,
    candidates are subjects (kge_model.py:164-181)."""
    if combine not in ("sp_", "_po"):
        raise ValueError('cannot handle combine="{}"'.format(combine))
    _require_cuda(q_tab, rel, cand_tab)
    lib, k = _lib.load(), _Keep()
    rq, rp, rc = k.rows(q_tab, q), k.rows(rel, p), k.rows(cand_tab, cand)
    n, m = int(rq.rows), int(rc.rows)
    if rp.rows != n:
        raise ValueError("need as many relation rows as entity rows")
    dev = q_tab.device
    if out is None:
        out = torch.empty((n, m), dtype=torch.float32, device=dev)
    ws = _workspace(MODELS[model], n, m, rq.dim, cand is not None, dev)
    _lib.check(lib.b200kge_score_1vsN(MODELS[model], SP_ if combine == "sp_" else _PO, l_norm,
                                      PREC[precision], C.byref(rq), C.byref(rp), C.byref(rc), n,
                                      out.data_ptr(), out.stride(0), ws.data_ptr(), ws.numel(),
                                      _stream(dev)))
    return out


def score_sp_po(model: str, ent, rel, s, p, o, entity_subset=None, l_norm: float = 1.0,
                precision: str = "auto", out: Optional[torch.Tensor] = None, ent_o=None, cand_tab=None):
    """[n, 2m] = [score_sp | score_po] in one launch sequence (kge_model.py:749-789).

    Index level: ent / rel are the tables, s / p / o index vectors.  Embedding level (s = p = o = None): ent, rel,
    ent_o are already-gathered [n, .] subject / relation / object rows and cand_tab the candidate table.
    `out` (optional) is a caller-owned [n, >= 2m] float32 block with unit column stride (e.g. a slice of an
    all-gather buffer): the two halves are written at columns [0, m) and [m, 2m)."""
    _require_cuda(ent, rel)
    lib, k = _lib.load(), _Keep()
    rs, rp, ro = k.rows(ent, s), k.rows(rel, p), k.rows(ent if ent_o is None else ent_o, o)
    rc = k.rows(ent if cand_tab is None else cand_tab, entity_subset)
    n, m = int(rs.rows), int(rc.rows)
    dev = ent.device
    if out is None:
        out = torch.empty((n, 2 * m), dtype=torch.float32, device=dev)
    elif out.dtype != torch.float32 or out.stride(1) != 1 or out.shape[0] < n or out.shape[1] < 2 * m:
        raise ValueError("out must be a float32 [n, >= 2m] block with unit column stride")
    ws = _workspace(MODELS[model], n, m, rs.dim, entity_subset is not None, dev)
    _lib.check(lib.b200kge_score_sp_po(MODELS[model], l_norm, PREC[precision], C.byref(rs), C.byref(rp),
                                       C.byref(ro), C.byref(rc), n, out.data_ptr(), out.stride(0),
                                       ws.data_ptr(), ws.numel(), _stream(dev)))
    return out


def score_sp_po_bcast(model: str, s_emb, rel, p, o_emb, cand_tab, out_ptr: int, peer_ptrs, ldo: int, col_block: int,
                      l_norm: float = 1.0, precision: str = "auto"):
    """score_sp_po whose epilogue stores go to this rank's buffer (raw device address out_ptr, already offset to the
    shard's first column) AND to the peer-mapped buffers `peer_ptrs` of the other ranks at the same offsets — the
    all-gather of per-shard logits fused into the scoring kernel (include/b200kge.h: b200kge_score_sp_po_bcast)."""
    _require_cuda(s_emb, rel, o_emb, cand_tab)
    lib, k = _lib.load(), _Keep()
    rs, rp, ro, rc = k.rows(s_emb), k.rows(rel, p), k.rows(o_emb), k.rows(cand_tab)
    n, m = int(rs.rows), int(rc.rows)
    dev = s_emb.device
    arr = (C.c_void_p * max(1, len(peer_ptrs)))(*[C.c_void_p(int(x)) for x in peer_ptrs])
    ws = _workspace(MODELS[model], n, m, rs.dim, False, dev)
    _lib.check(lib.b200kge_score_sp_po_bcast(MODELS[model], l_norm, PREC[precision], C.byref(rs), C.byref(rp),
                                             C.byref(ro), C.byref(rc), n, C.c_void_p(int(out_ptr)), arr,
                                             len(peer_ptrs), ldo, col_block, ws.data_ptr(), ws.numel(), _stream(dev)))


def rank_sp_po(model: str, s_tab, rel, o_tab, cand_tab, true_scores, s=None, p=None, o=None, cand=None,
               filter_labels=None, rtol: float = 1e-4, atol: float = 1e-5, l_norm: float = 1.0,
               precision: str = "auto", rank=None, ties=None):
    """Both directions of one batch's ranking against one chunk of candidates in ONE launch sequence: rows
    0..n-1 = sp_ queries, n..2n-1 = _po queries; true_scores / rank / ties are [2n] in that order (rank / ties
    int64, accumulated into); filter_labels (optional) [2n, m]."""
    _require_cuda(s_tab, rel, o_tab, cand_tab, true_scores, filter_labels)
    lib, k = _lib.load(), _Keep()
    rs, rp, ro, rc = k.rows(s_tab, s), k.rows(rel, p), k.rows(o_tab, o), k.rows(cand_tab, cand)
    n, m = int(rs.rows), int(rc.rows)
    dev = s_tab.device
    t = true_scores.reshape(-1).float().contiguous()
    if t.numel() != 2 * n:
        raise ValueError("true_scores must hold 2n values: sp_ rows first, then _po rows")
    if rank is None:
        rank = torch.zeros(2 * n, dtype=torch.int64, device=dev)
    if ties is None:
        ties = torch.zeros(2 * n, dtype=torch.int64, device=dev)
    f = None
    if filter_labels is not None:
        f = filter_labels if (filter_labels.dtype == torch.float32 and filter_labels.stride(1) == 1) \
            else filter_labels.float().contiguous()
    ws = _workspace(MODELS[model], n, m, rs.dim, cand is not None, dev)
    _lib.check(lib.b200kge_rank_sp_po(
        MODELS[model], l_norm, PREC[precision], C.byref(rs), C.byref(rp), C.byref(ro), C.byref(rc), n, t.data_ptr(),
        f.data_ptr() if f is not None else None, f.stride(0) if f is not None else 0, rtol, atol, rank.data_ptr(),
        ties.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)))
    return rank, ties


def rank_sp_po_csr(model: str, s_tab, rel, o_tab, cand_tab, true_scores, filter_off, filter_col, own_col=None,
                   s=None, p=None, o=None, rtol: float = 1e-4, atol: float = 1e-5, l_norm: float = 1.0,
                   precision: str = "auto", rank=None, ties=None):
    """rank_sp_po with the known-answer filter as CSR over the stacked rows (sp_ rows, then _po rows): row r lists
    sorted candidate columns filter_col[filter_off[r]:filter_off[r+1]]; own_col[r] (the row's own answer) stays in.
    Raises NotImplementedError when the shape / model is not served by the pre-split tensor-core kernel."""
    _require_cuda(s_tab, rel, o_tab, cand_tab, true_scores, filter_off, filter_col, own_col)
    lib, k = _lib.load(), _Keep()
    rs, rp, ro, rc = k.rows(s_tab, s), k.rows(rel, p), k.rows(o_tab, o), k.rows(cand_tab)
    n, m = int(rs.rows), int(rc.rows)
    dev = s_tab.device
    t = true_scores.reshape(-1).float().contiguous()
    offs, cols = _i64(filter_off), _i64(filter_col)
    own = _i64(own_col)
    if t.numel() != 2 * n or offs.numel() != 2 * n + 1:
        raise ValueError("true_scores / filter_off must cover the 2n stacked rows")
    if rank is None:
        rank = torch.zeros(2 * n, dtype=torch.int64, device=dev)
    if ties is None:
        ties = torch.zeros(2 * n, dtype=torch.int64, device=dev)
    ws = _workspace(MODELS[model], n, m, rs.dim, False, dev)
    _lib.check(lib.b200kge_rank_sp_po_csr(
        MODELS[model], l_norm, PREC[precision], C.byref(rs), C.byref(rp), C.byref(ro), C.byref(rc), n, t.data_ptr(),
        offs.data_ptr(), cols.data_ptr() if cols.numel() else None, own.data_ptr() if own is not None else None,
        rtol, atol, rank.data_ptr(), ties.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)))
    return rank, ties


def shard_gather_rows(shard: torch.Tensor, lo: int, idx: torch.Tensor, out: Optional[torch.Tensor] = None):
    """This rank's contribution to the query-row exchange of an entity-sharded table: out[i] = shard[idx[i] - lo]
    if lo <= idx[i] < lo + rows else 0 (one kernel, no host synchronisation)."""
    _require_cuda(shard, idx)
    lib, k = _lib.load(), _Keep()
    rsh = k.rows(shard)
    ix = _i64(idx)
    n = ix.numel()
    if out is None:
        out = torch.empty((n, shard.shape[1]), dtype=torch.float32, device=shard.device)
    _lib.check(lib.b200kge_shard_gather_rows(C.byref(rsh), int(lo), ix.data_ptr(), n, out.data_ptr(), out.stride(0),
                                             _stream(shard.device)))
    return out


def _labels(k: _Keep, labels: torch.Tensor) -> Labels:
    lab = Labels()
    if labels.dim() == 1:
        li = _i64(labels)
        k.refs.append(li)
        lab.idx, lab.dense, lab.ldl = li.data_ptr(), None, 0
    else:
        ld = labels if (labels.dtype == torch.float32 and labels.stride(1) == 1) else labels.float().contiguous()
        k.refs.append(ld)
        lab.idx, lab.dense, lab.ldl = None, ld.data_ptr(), ld.stride(0)
    return lab


def score_1vsN_loss(model: str, combine: str, q_tab, rel, cand_tab, labels, q=None, p=None, cand=None,
                    loss: str = "bce", offset: float = 0.0, l_norm: float = 1.0, precision: str = "auto",
                    return_rows: bool = False):
    """Fused scoring + KgeLoss (sum reduction): returns a 0-d tensor (and per-row terms)."""
    _require_cuda(q_tab, rel, cand_tab, labels)
    lib, k = _lib.load(), _Keep()
    rq, rp, rc = k.rows(q_tab, q), k.rows(rel, p), k.rows(cand_tab, cand)
    n, m = int(rq.rows), int(rc.rows)
    dev = q_tab.device
    lab = _labels(k, labels)
    out = torch.empty((), dtype=torch.float32, device=dev)
    rows = torch.empty(n, dtype=torch.float32, device=dev) if return_rows else None
    ws = _workspace(MODELS[model], n, m, rq.dim, cand is not None, dev)
    _lib.check(lib.b200kge_score_1vsN_loss(
        MODELS[model], SP_ if combine == "sp_" else _PO, l_norm, PREC[precision], C.byref(rq), C.byref(rp),
        C.byref(rc), n, C.byref(lab), LOSS[loss], offset, out.data_ptr(),
        rows.data_ptr() if rows is not None else None, ws.data_ptr(), ws.numel(), _stream(dev)))
    return (out, rows) if return_rows else out


def score_1vsN_rank(model: str, combine: str, q_tab, rel, cand_tab, true_scores, q=None, p=None, cand=None,
                    filter_labels=None, rtol: float = 1e-4, atol: float = 1e-5, l_norm: float = 1.0,
                    precision: str = "auto", rank=None, ties=None):
    """Fused scoring + rank/tie counting for one chunk of candidates; accumulates into rank/ties."""
    _require_cuda(q_tab, rel, cand_tab, true_scores, filter_labels)
    lib, k = _lib.load(), _Keep()
    rq, rp, rc = k.rows(q_tab, q), k.rows(rel, p), k.rows(cand_tab, cand)
    n, m = int(rq.rows), int(rc.rows)
    dev = q_tab.device
    t = true_scores.reshape(-1).float().contiguous()
    if rank is None:
        rank = torch.zeros(n, dtype=torch.int64, device=dev)
    if ties is None:
        ties = torch.zeros(n, dtype=torch.int64, device=dev)
    f = None
    if filter_labels is not None:
        f = filter_labels if (filter_labels.dtype == torch.float32 and filter_labels.stride(1) == 1) \
            else filter_labels.float().contiguous()
    ws = _workspace(MODELS[model], n, m, rq.dim, cand is not None, dev)
    _lib.check(lib.b200kge_score_1vsN_rank(
        MODELS[model], SP_ if combine == "sp_" else _PO, l_norm, PREC[precision], C.byref(rq), C.byref(rp),
        C.byref(rc), n, t.data_ptr(), f.data_ptr() if f is not None else None,
        f.stride(0) if f is not None else 0, rtol, atol, rank.data_ptr(), ties.data_ptr(), ws.data_ptr(),
        ws.numel(), _stream(dev)))
    return rank, ties


def loss_dense(scores, labels, loss: str = "bce", offset: float = 0.0, return_rows: bool = False):
    """KgeLoss (sum) on a dense score matrix (loss.py:153-159 / :198-213)."""
    _require_cuda(scores, labels)
    lib, k = _lib.load(), _Keep()
    x = scores if (scores.dtype == torch.float32 and scores.stride(1) == 1) else scores.float().contiguous()
    n, m = x.shape
    lab = _labels(k, labels)
    dev = x.device
    out = torch.empty((), dtype=torch.float32, device=dev)
    rows = torch.empty(n, dtype=torch.float32, device=dev) if return_rows else None
    nbytes = n * ((m + 4095) // 4096) * 5 * 4 + 4096
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _lib.check(lib.b200kge_loss_dense(x.data_ptr(), x.stride(0), n, m, C.byref(lab), LOSS[loss], offset,
                                      out.data_ptr(), rows.data_ptr() if rows is not None else None,
                                      ws.data_ptr(), ws.numel(), _stream(dev)))
    return (out, rows) if return_rows else out


def rank_dense(scores, true_scores, filter_labels=None, rtol: float = 1e-4, atol: float = 1e-5,
               rank=None, ties=None):
    """_get_ranks_and_num_ties (+ optional filter subtraction) on dense scores; integer, bit-exact."""
    _require_cuda(scores, true_scores, filter_labels)
    lib = _lib.load()
    x = scores if (scores.dtype == torch.float32 and scores.stride(1) == 1) else scores.float().contiguous()
    n, m = x.shape
    dev = x.device
    t = true_scores.reshape(-1).float().contiguous()
    if rank is None:
        rank = torch.zeros(n, dtype=torch.int64, device=dev)
    if ties is None:
        ties = torch.zeros(n, dtype=torch.int64, device=dev)
    f = None
    if filter_labels is not None:
        f = filter_labels if (filter_labels.dtype == torch.float32 and filter_labels.stride(1) == 1) \
            else filter_labels.float().contiguous()
    _lib.check(lib.b200kge_rank_dense(x.data_ptr(), x.stride(0), n, m, t.data_ptr(),
                                      f.data_ptr() if f is not None else None,
                                      f.stride(0) if f is not None else 0, rtol, atol, rank.data_ptr(),
                                      ties.data_ptr(), _stream(dev)))
    return rank, ties


def ns_score(model: str, ent, rel, triples, negatives, slot: int, with_positive: bool = False,
             l_norm: float = 1.0):
    """[n, K] (or [n, 1+K] with the positive in column 0) negative-sample scores."""
    _require_cuda(ent, rel, triples, negatives)
    lib, k = _lib.load(), _Keep()
    tri = triples.long()
    rs, rp, ro = k.rows(ent, tri[:, S].contiguous()), k.rows(rel, tri[:, P].contiguous()), \
        k.rows(ent, tri[:, O].contiguous())
    neg = negatives.long().contiguous()
    n, K = neg.shape
    table = k.rows(rel if slot == P else ent)
    dev = ent.device
    out = torch.empty((n, K + (1 if with_positive else 0)), dtype=torch.float32, device=dev)
    _lib.check(lib.b200kge_ns_score(MODELS[model], l_norm, C.byref(rs), C.byref(rp), C.byref(ro),
                                    C.byref(table), slot, neg.data_ptr(), n, K, 1 if with_positive else 0,
                                    out.data_ptr(), out.stride(0), _stream(dev)))
    return out


def sample_uniform(n: int, K: int, vocab: int, seed: int, offset: int, device) -> torch.Tensor:
    """[n, K] int64 ids ~ U{0..vocab-1} drawn on the device (Philox4x32-10 keyed by seed, counter = (position, offset))."""
    lib = _lib.load()
    out = torch.empty((n, K), dtype=torch.int64, device=device)
    if not out.is_cuda:
        raise RuntimeError("kge_b200 runs on CUDA (sm_100) tensors only; there is no CPU path")
    _lib.check(lib.b200kge_sample_uniform(seed & (2 ** 64 - 1), offset & (2 ** 64 - 1), vocab, n, K, out.data_ptr(),
                                          _stream(out.device)))
    return out


def train_1vsall_forward(model: str, ent, rel, triples, loss: str = "bce", offset: float = 0.0,
                         l_norm: float = 1.0, precision: str = "auto", out=None, workspace=None):
    """One fused 1vsAll forward step for device-resident triples [n,3]; returns the 0-d loss
    (loss(score_sp,o) + loss(score_po,s)) / n  (train_1vsAll.py:48-82)."""
    _require_cuda(ent, rel, triples)
    lib, k = _lib.load(), _Keep()
    re_, rr = k.rows(ent), k.rows(rel)
    tri = triples if (triples.dtype == torch.int64 and triples.is_contiguous()) else triples.long().contiguous()
    n = tri.shape[0]
    dev = ent.device
    if out is None:
        out = torch.empty((), dtype=torch.float32, device=dev)
    ws = workspace if workspace is not None else _workspace(MODELS[model], n, ent.shape[0], ent.shape[1], False, dev)
    _lib.check(lib.b200kge_train_1vsall_forward(
        MODELS[model], l_norm, PREC[precision], C.byref(re_), C.byref(rr), tri.data_ptr(), n, LOSS[loss],
        offset, out.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)))
    return out


class Step1vsAll:
    """A prepared fused 1vsAll forward step for device-resident batches: table views, enums, workspace and the
    output scalar are set up once, a call is ONE ctypes call (the job plugin's per-batch path; saves ~20 us of Python
    per step against train_1vsall_forward).  The returned 0-d tensor is a persistent buffer that the next call
    overwrites — read it (`.item()`) before calling again.  Valid as long as the tables keep their storage."""

    def __init__(self, model: str, ent: torch.Tensor, rel: torch.Tensor, max_n: int, loss: str = "bce",
                 offset: float = 0.0, l_norm: float = 1.0, precision: str = "auto"):
        _require_cuda(ent, rel)
        self.lib = _lib.load()
        self.ent, self.rel = _f32(ent), _f32(rel)
        self.k = _Keep()
        self.re, self.rr = self.k.rows(self.ent), self.k.rows(self.rel)
        self.max_n = int(max_n)
        nbytes = self.lib.b200kge_workspace_bytes(MODELS[model], self.max_n, ent.shape[0], ent.shape[1], 0)
        self.ws = torch.empty(nbytes + self.max_n * 24 + 1024, dtype=torch.uint8, device=ent.device)   # + staged batch (host form)
        self.out = torch.zeros((), dtype=torch.float32, device=ent.device)
        self.loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
        self.loss_np = self.loss_host.numpy()
        self.key = (self.ent.data_ptr(), self.rel.data_ptr(), tuple(ent.shape), tuple(rel.shape))
        self.args = (MODELS[model], C.c_float(l_norm), PREC[precision], C.byref(self.re), C.byref(self.rr))
        self.tail = (LOSS[loss], C.c_float(offset), C.c_void_p(self.out.data_ptr()), C.c_void_p(self.ws.data_ptr()),
                     self.ws.numel())
        self.dev = ent.device

    def matches(self, ent, rel, n):
        return n <= self.max_n and self.key == (ent.data_ptr(), rel.data_ptr(), tuple(ent.shape), tuple(rel.shape))

    def __call__(self, triples: torch.Tensor) -> torch.Tensor:
        if triples.dtype != torch.int64 or not triples.is_contiguous():
            triples = triples.long().contiguous()
        rc = self.lib.b200kge_train_1vsall_forward(*self.args, C.c_void_p(triples.data_ptr()), triples.shape[0],
                                                   *self.tail, _stream(self.dev))
        if rc:
            _lib.check(rc)
        return self.out

    def call_host(self, triples_host: torch.Tensor) -> float:
        """The same step from a HOST batch (contiguous int64 [n,3], ideally pinned): host->device copy, kernels, the
        4-byte read-back and the stream synchronisation inside ONE library call (triples.to(device) ... .item(),
        train_1vsAll.py:59-77)."""
        rc = self.lib.b200kge_train_1vsall_forward_host(
            *self.args, C.c_void_p(triples_host.data_ptr()), triples_host.shape[0], self.tail[0], self.tail[1],
            C.c_void_p(self.loss_host.data_ptr()), self.tail[3], self.ws.numel(), _stream(self.dev))
        if rc:
            _lib.check(rc)
        return float(self.loss_np[0])


class HostStep:
    """End-to-end 1vsAll forward step with HOST buffers (pinned in, scalar out): the call a
    training loop makes per batch — triples.to(device) ... loss.item() (train_1vsAll.py:59-77)."""

    def __init__(self, model: str, ent: torch.Tensor, rel: torch.Tensor, max_n: int, loss: str = "bce",
                 offset: float = 0.0, l_norm: float = 1.0, precision: str = "auto"):
        _require_cuda(ent, rel)
        self.model, self.loss, self.offset, self.l_norm, self.precision = model, loss, offset, l_norm, precision
        self.ent, self.rel = _f32(ent), _f32(rel)
        self.k = _Keep()
        self.re, self.rr = self.k.rows(self.ent), self.k.rows(self.rel)
        self.ws = _workspace(MODELS[model], max_n, ent.shape[0], ent.shape[1], False, ent.device)
        self.loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
        self.h2d_bytes = 0
        self.d2h_bytes = 4

    def __call__(self, triples_host: torch.Tensor) -> float:
        if triples_host.is_cuda or triples_host.dtype != torch.int64 or not triples_host.is_contiguous():
            raise ValueError("triples_host must be a contiguous int64 CPU tensor [n,3]")
        n = triples_host.shape[0]
        self.h2d_bytes = n * 3 * 8
        _lib.check(_lib.load().b200kge_train_1vsall_forward_host(
            MODELS[self.model], self.l_norm, PREC[self.precision], C.byref(self.re), C.byref(self.rr),
            triples_host.data_ptr(), n, LOSS[self.loss], self.offset, self.loss_host.data_ptr(),
            self.ws.data_ptr(), self.ws.numel(), _stream(self.ent.device)))
        return float(self.loss_host[0])


# ---- SURVEY 8f rows (gradients, penalties, CSR labels): validated on a B200 in round 2 -------------------------
def gemm_nt(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """C = A @ B^T (fp32 in/out) on the f16 tensor pipe from pre-split hi/lo fp16 planes."""
    _require_cuda(a, b)
    lib = _lib.load()
    a, b = _f32(a), _f32(b)
    M, K = a.shape
    N = b.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    ws = torch.empty(lib.b200kge_gemm_nt_workspace_bytes(M, N, K), dtype=torch.uint8, device=a.device)
    _lib.check(lib.b200kge_gemm_nt(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), M, N, K, out.data_ptr(),
                                     out.stride(0), ws.data_ptr(), ws.numel(), _stream(a.device)))
    return out


def train_1vsall_backward(model: str, ent, rel, triples, loss: str = "bce", offset: float = 0.0, l_norm: float = 1.0):
    """(d_ent, d_rel): dense table gradients of train_1vsall_forward's loss (dot family; TransE L1/L2; RotatE L1)."""
    _require_cuda(ent, rel, triples)
    lib, k = _lib.load(), _Keep()
    re_, rr = k.rows(ent), k.rows(rel)
    tri = triples if (triples.dtype == torch.int64 and triples.is_contiguous()) else triples.long().contiguous()
    n = tri.shape[0]
    dev = ent.device
    d_ent = torch.empty_like(_f32(ent))
    d_rel = torch.empty_like(_f32(rel))
    nbytes = lib.b200kge_train_1vsall_backward_workspace_bytes(MODELS[model], n, ent.shape[0], ent.shape[1])
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _lib.check(lib.b200kge_train_1vsall_backward(
        MODELS[model], l_norm, C.byref(re_), C.byref(rr), tri.data_ptr(), n, LOSS[loss], offset, d_ent.data_ptr(),
        d_ent.stride(0), d_rel.data_ptr(), d_rel.stride(0), ws.data_ptr(), ws.numel(), _stream(dev)))
    return d_ent, d_rel


def score_1vsN_backward(model: str, combine: str, ent, rel, q, p, grad_scores, l_norm: float = 1.0):
    """(d_ent, d_rel) of a dense [n, E] score block given dL/dscores (fresh tensors): dot family (tensor cores), TransE
    L1 / L2 and RotatE L1 (row-gradient kernel)."""
    _require_cuda(ent, rel, grad_scores)
    lib, k = _lib.load(), _Keep()
    re_, rr = k.rows(ent), k.rows(rel)
    qi, pi = _i64(q), _i64(p)
    g = grad_scores if (grad_scores.dtype == torch.float32 and grad_scores.stride(1) == 1) else grad_scores.float().contiguous()
    n = qi.numel()
    dev = ent.device
    d_ent = torch.empty_like(_f32(ent))
    d_rel = torch.empty_like(_f32(rel))
    ws = torch.empty(lib.b200kge_score_1vsN_backward_workspace_bytes(MODELS[model], n, ent.shape[0], ent.shape[1]),
                     dtype=torch.uint8, device=dev)
    _lib.check(lib.b200kge_score_1vsN_backward(
        MODELS[model], SP_ if combine == "sp_" else _PO, l_norm, C.byref(re_), C.byref(rr), qi.data_ptr(), pi.data_ptr(), n,
        g.data_ptr(), g.stride(0), d_ent.data_ptr(), d_ent.stride(0), d_rel.data_ptr(), d_rel.stride(0), ws.data_ptr(),
        ws.numel(), _stream(dev)))
    return d_ent, d_rel


def score_1vsN_loss_csr_backward(model: str, combine: str, ent, rel, q, p, csr_offsets, csr_cols, loss: str = "kl",
                                 offset: float = 0.0, label_smoothing: float = 0.0, batch_size: Optional[int] = None):
    """(d_ent, d_rel) of score_1vsN_loss_csr(...) / batch_size over the whole entity table (dot family)."""
    _require_cuda(ent, rel, csr_offsets, csr_cols)
    lib, k = _lib.load(), _Keep()
    re_, rr = k.rows(ent), k.rows(rel)
    qi, pi, offs, cols = _i64(q), _i64(p), _i64(csr_offsets), _i64(csr_cols)
    n = qi.numel()
    dev = ent.device
    d_ent = torch.empty_like(_f32(ent))
    d_rel = torch.empty_like(_f32(rel))
    ws = torch.empty(lib.b200kge_score_1vsN_backward_workspace_bytes(MODELS[model], n, ent.shape[0], ent.shape[1]),
                     dtype=torch.uint8, device=dev)
    _lib.check(lib.b200kge_score_1vsN_loss_csr_backward(
        MODELS[model], SP_ if combine == "sp_" else _PO, C.byref(re_), C.byref(rr), qi.data_ptr(), pi.data_ptr(), n,
        offs.data_ptr(), cols.data_ptr() if cols.numel() else None, label_smoothing, LOSS[loss], offset,
        batch_size or n, d_ent.data_ptr(), d_ent.stride(0), d_rel.data_ptr(), d_rel.stride(0), ws.data_ptr(), ws.numel(),
        _stream(dev)))
    return d_ent, d_rel


def lookup_penalty(weight: torch.Tensor, regularize: str = "lp", regularize_weight: float = 0.0, p: float = 2.0,
                     weighted: bool = False, indexes: Optional[torch.Tensor] = None, space: str = "euclidean"):
    """LookupEmbedder.penalty (lookup_embedder.py:123-177) as a 0-d tensor."""
    _require_cuda(weight, indexes)
    dev = weight.device
    if regularize == "" or regularize_weight == 0.0:
        return torch.zeros((), dtype=torch.float32, device=dev)
    if regularize == "n3":
        p = 3.0
    elif regularize != "lp":
        raise ValueError(f"Invalid value regularize={regularize}")
    lib, k = _lib.load(), _Keep()
    if regularize == "n3" and space != "complex":
        # the reference accepts n3 only in complex space (lookup_embedder.py:29-34), so its signed-cube branch for
        # weighted n3 elsewhere (:158) is unreachable
        raise ValueError("Illegal value n3 for key regularize; allowed values are ['', 'lp'] (space is not complex)")
    complex_abs = 1 if regularize == "n3" else 0
    counts = None
    if weighted:
        uniq, cnt = torch.unique(indexes, return_counts=True)
        rows = k.rows(weight, uniq)
        counts = cnt.float().contiguous()
        scale = regularize_weight / p / indexes.shape[0]          # len(indexes): rows of the index block
    else:
        rows = k.rows(weight)
        scale = regularize_weight / p
    out = torch.empty((), dtype=torch.float32, device=dev)
    ws = torch.empty(((int(rows.rows) + 7) // 8 + 2) * 4, dtype=torch.uint8, device=dev)
    _lib.check(lib.b200kge_lookup_penalty(C.byref(rows), counts.data_ptr() if counts is not None else None, p,
                                            complex_abs, scale, out.data_ptr(), ws.data_ptr(), ws.numel(), _stream(dev)))
    return out


def normalize_rows_(weight: torch.Tensor, p: float) -> torch.Tensor:
    """In-place row normalisation to unit Lp norm (lookup_embedder.py:64-69)."""
    _require_cuda(weight)
    w = _f32(weight)
    if w.data_ptr() != weight.data_ptr():
        raise ValueError("normalisation is in place: pass a row-contiguous float32 matrix")
    _lib.check(_lib.load().b200kge_normalize_rows(w.data_ptr(), w.stride(0), w.shape[0], w.shape[1], p,
                                                    _stream(w.device)))
    return weight


def ns_backward(model: str, ent, rel, triples, negatives: dict, offset: float = 0.0, l_norm: float = 1.0,
                  batch_size: Optional[int] = None):
    """(d_ent, d_rel) of one negative-sampling batch with BCE; negatives = {slot: [n, K] ids}, slots 0 (S), 2 (O)."""
    _require_cuda(ent, rel, triples)
    lib, k = _lib.load(), _Keep()
    re_, rr = k.rows(ent), k.rows(rel)
    tri = triples if (triples.dtype == torch.int64 and triples.is_contiguous()) else triples.long().contiguous()
    n = tri.shape[0]
    dev = ent.device
    d_ent = torch.zeros_like(_f32(ent))
    d_rel = torch.zeros_like(_f32(rel))
    ws = torch.empty(n * (ent.shape[1] + 32) * 4 + 1024, dtype=torch.uint8, device=dev)
    for slot, neg in negatives.items():
        ng = neg if (neg.dtype == torch.int64 and neg.is_contiguous()) else neg.long().contiguous()
        _lib.check(lib.b200kge_ns_backward(
            MODELS[model], l_norm, C.byref(re_), C.byref(rr), tri.data_ptr(), int(slot), ng.data_ptr(), n, ng.shape[1],
            offset, batch_size or n, d_ent.data_ptr(), d_ent.stride(0), d_rel.data_ptr(), d_rel.stride(0),
            ws.data_ptr(), ws.numel(), _stream(dev)))
    return d_ent, d_rel


def score_1vsN_loss_csr(model: str, combine: str, q_tab, rel, cand_tab, csr_offsets, csr_cols, q=None, p=None,
                          loss: str = "kl", offset: float = 0.0, label_smoothing: float = 0.0, l_norm: float = 1.0,
                          precision: str = "auto", return_rows: bool = False):
    """KvsAll loss (sum over rows) with CSR multi-hot labels — see b200kge_score_1vsN_loss_csr."""
    _require_cuda(q_tab, rel, cand_tab, csr_offsets, csr_cols)
    lib, k = _lib.load(), _Keep()
    rq, rp, rc = k.rows(q_tab, q), k.rows(rel, p), k.rows(cand_tab)
    n, m = int(rq.rows), int(rc.rows)
    dev = q_tab.device
    offs, cols = _i64(csr_offsets), _i64(csr_cols)
    nnz = int(cols.numel())
    out = torch.empty((), dtype=torch.float32, device=dev)
    rows = torch.empty(n, dtype=torch.float32, device=dev) if return_rows else None
    nbytes = lib.b200kge_score_1vsN_loss_csr_workspace_bytes(MODELS[model], n, m, rq.dim, nnz)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _lib.check(lib.b200kge_score_1vsN_loss_csr(
        MODELS[model], SP_ if combine == "sp_" else _PO, l_norm, PREC[precision], C.byref(rq), C.byref(rp), C.byref(rc),
        n, offs.data_ptr(), cols.data_ptr() if nnz else None, nnz, label_smoothing, LOSS[loss], offset, out.data_ptr(),
        rows.data_ptr() if rows is not None else None, ws.data_ptr(), ws.numel(), _stream(dev)))
    return (out, rows) if return_rows else out
