"""Entity-sharded scoring across GPUs (one process per GPU, torch.distributed / NCCL over NVLink).

The reference has no distributed path; its natural seam is the entity-chunk loop of
entity ranking, whose per-chunk rank/tie counts are additive (eval_entity_ranking.py:222-229,
310-313).  For graphs whose entity table outgrows one GPU (Wikidata5M: 4.8 M x 512 fp32 = 9.8 GB,
plus optimizer state) rank g owns rows [g*ceil(E/G), (g+1)*ceil(E/G)); the relation table is
replicated (<= 6 MB).  One call =

  1. query-row exchange: each rank fills the s/o rows it owns into a zeroed [n, D] buffer
     -> all-reduce(sum)                                        (2 * n * D * 4 bytes)
  2. local-shard scoring with the same single-GPU kernels     ([n, E/G] per direction, fused epilogues)
  3. result exchange, by consumer:
       score_sp_po   -> all-gather of per-shard logits         (north_star)
       rank_sp_po    -> all-reduce(sum) of int64 rank/tie counts  (integer => bit-exact, 4*n*8 bytes)
       loss_1vsall   -> all-reduce(sum) of per-rank BCE partial sums
       topk_sp       -> local top-k, all-gather of [n,k] (value,index), merge (lowest index wins ties)

Small tables (FB15k-237, YAGO3-10) are better served by replicas + batch split (no collective);
that is what bench.py --gpus N runs.

The local scorer is pluggable: `EngineBackend` (CUDA kernels through the C ABI) in production;
the CPU gloo tests inject an oracle-based backend to exercise the host logic without a GPU.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


class EngineBackend:
    """Local-shard scoring on the GPU through libb200kge."""

    def exchange_rows(self, shard, lo, idx):
        from . import engine
        return engine.shard_gather_rows(shard, lo, idx)

    def score_sp_po(self, model, s_emb, rel, p, o_emb, cand, l_norm, precision, out=None):
        from . import engine
        return engine.score_sp_po(model, s_emb, rel, None, p, None, None, l_norm, precision, out=out, ent_o=o_emb,
                                  cand_tab=cand)

    def rank_sp_po(self, model, s_emb, rel, p, o_emb, cand, true2n, filter2n, rtol, atol, l_norm, precision):
        from . import engine
        if isinstance(filter2n, tuple):          # CSR filter (offsets [2n+1], local columns, own local column [2n])
            off, col, own = filter2n
            return engine.rank_sp_po_csr(model, s_emb, rel, o_emb, cand, true2n, off, col, own, None, p, None, rtol, atol,
                                         l_norm, precision)
        return engine.rank_sp_po(model, s_emb, rel, o_emb, cand, true2n, None, p, None, None, filter2n, rtol, atol,
                                 l_norm, precision)

    def score_1vsN(self, model, combine, q_emb, p_emb, cand, l_norm, precision):
        from . import engine
        return engine.score_1vsN(model, combine, q_emb, p_emb, cand, l_norm=l_norm, precision=precision)

    def score_spo(self, model, s_emb, p_emb, o_emb, l_norm):
        from . import engine
        return engine.score_spo(model, s_emb, p_emb, o_emb, l_norm=l_norm)

    def rank_1vsN(self, model, combine, q_emb, p_emb, cand, true_scores, filter_labels, rtol, atol, l_norm,
                  precision):
        from . import engine
        return engine.score_1vsN_rank(model, combine, q_emb, p_emb, cand, true_scores, None, None, None,
                                      filter_labels, rtol, atol, l_norm, precision)

    def bce_1vsN(self, model, combine, q_emb, p_emb, cand, local_labels, offset, l_norm, precision):
        from . import engine
        return engine.score_1vsN_loss(model, combine, q_emb, p_emb, cand, local_labels, None, None, None,
                                      "bce", offset, l_norm, precision)


class ShardedKgeModel:
    def __init__(self, model: str, ent_shard: torch.Tensor, rel: torch.Tensor, num_entities: int,
                 rank: Optional[int] = None, world: Optional[int] = None, group=None,
                 l_norm: float = 1.0, precision: str = "auto", backend=None):
        self.model, self.l_norm, self.precision = model, l_norm, precision
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.E = int(num_entities)
        self.per = (self.E + self.world - 1) // self.world
        self.lo = min(self.rank * self.per, self.E)
        self.hi = min(self.lo + self.per, self.E)
        if ent_shard.shape[0] != self.hi - self.lo:
            raise ValueError(f"rank {self.rank} must hold rows [{self.lo},{self.hi}) of the entity table, "
                             f"got {ent_shard.shape[0]} rows")
        self.ent, self.rel = ent_shard, rel
        self.backend = backend or EngineBackend()

    # -- ownership ---------------------------------------------------------------------------------
    @staticmethod
    def shard_bounds(num_entities: int, world: int, rank: int):
        per = (num_entities + world - 1) // world
        lo = min(rank * per, num_entities)
        return lo, min(lo + per, num_entities)

    def _all_reduce(self, t, op=dist.ReduceOp.SUM):
        if self.world > 1:
            dist.all_reduce(t, op=op, group=self.group)
        return t

    # -- step 1: query-row exchange ------------------------------------------------------------------
    def gather_entity_rows(self, idx: torch.Tensor) -> torch.Tensor:
        """[n, D] rows of the GLOBAL entity table for global ids `idx`, on every rank: each rank writes the rows it
        owns (zeros elsewhere; one kernel, no host synchronisation), then one all-reduce(sum) — every other rank
        contributed exact zeros, so the sum is a copy."""
        out = self.backend.exchange_rows(self.ent, self.lo, idx.long())
        return self._all_reduce(out)

    def _queries(self, s, p, o):
        both = self.gather_entity_rows(torch.cat([s.long(), o.long()]))
        n = s.numel()
        return both[:n], p.long(), both[n:], both

    # -- full logits ---------------------------------------------------------------------------------
    def score_sp_po(self, s, p, o) -> torch.Tensor:
        """[n, 2E] = [score_sp | score_po] (kge_model.py:749-789) assembled from per-shard logits: both directions
        of the shard are scored by ONE stacked launch straight into this rank's slot of the all-gather buffer."""
        s_emb, pi, o_emb, _ = self._queries(s, p, o)
        n = s.numel()
        m = self.hi - self.lo
        dev = self.ent.device
        buf = torch.empty((self.world, n, 2 * self.per), dtype=torch.float32, device=dev)
        mine = buf[self.rank]
        if m == self.per:
            self.backend.score_sp_po(self.model, s_emb, self.rel, pi, o_emb, self.ent, self.l_norm, self.precision,
                                     out=mine)
        else:                                     # ragged last shard: pad its columns
            mine.zero_()
            if m > 0:
                loc = self.backend.score_sp_po(self.model, s_emb, self.rel, pi, o_emb, self.ent, self.l_norm,
                                               self.precision)
                mine[:, :m] = loc[:, :m]
                mine[:, self.per:self.per + m] = loc[:, m:]
        if self.world > 1:
            dist.all_gather_into_tensor(buf.view(-1), mine.reshape(-1), group=self.group)
        # [world, n, 2, per] -> [n, 2, world * per] -> [n, 2E]: the one re-layout copy of the call
        full = buf.view(self.world, n, 2, self.per).permute(1, 2, 0, 3).reshape(n, 2, self.world * self.per)
        return full[:, :, : self.E].reshape(n, 2 * self.E)

    # -- full logits, compute + collective in one kernel ---------------------------------------------------
    def _symm_logits(self, n):
        """A symmetric [n, 2E] buffer (torch.distributed._symmetric_memory: same allocation on every rank, mapped
        into every peer's address space over NVLink) and the peers' device pointers to it; cached per n."""
        cache = self.__dict__.setdefault("_symm_cache", {})
        if n not in cache:
            import torch.distributed._symmetric_memory as symm

            buf = symm.empty((n, 2 * self.E), dtype=torch.float32, device=self.ent.device)
            hdl = symm.rendezvous(buf, self.group if self.group is not None else dist.group.WORLD)
            cache[n] = (buf, hdl, [int(x) for x in hdl.buffer_ptrs])
        return cache[n]

    def score_sp_po_fused(self, s, p, o) -> torch.Tensor:
        """[n, 2E] logits like score_sp_po, with the all-gather FUSED into the scoring kernel: its epilogue stores
        every score of the shard straight into the [n, 2E] matrices of all ranks (peer-mapped symmetric memory over
        NVLink / NVSwitch) at their final positions — no NCCL all-gather, no re-layout copy, the transfer overlaps
        the scoring tile by tile.  The returned tensor is the cached symmetric buffer (overwritten by the next call
        with the same n)."""
        s_emb, pi, o_emb, _ = self._queries(s, p, o)
        n = s.numel()
        if self.world == 1:
            return self.score_sp_po(s, p, o)
        buf, hdl, ptrs = self._symm_logits(n)
        hdl.barrier(channel=0)                  # every rank is done with the previous contents
        off = self.lo * 4
        if self.hi > self.lo:
            from . import engine
            engine.score_sp_po_bcast(self.model, s_emb, self.rel, pi, o_emb, self.ent, ptrs[self.rank] + off,
                                     [ptrs[g] + off for g in range(self.world) if g != self.rank], 2 * self.E,
                                     self.E, self.l_norm, self.precision)
        hdl.barrier(channel=1)                  # all shards have landed everywhere
        return buf

    # -- ranking -------------------------------------------------------------------------------------
    def true_scores(self, s, p, o):
        """((t_sp [n], t_po [n]), rows): scores of the true triples computed WITH THE 1-vs-N CODE PATH against the
        exchanged answer rows — the reference does the same to keep tie handling consistent
        (eval_entity_ranking.py:184-203: "scoring with spo vs sp and po can lead to slight differences for ties");
        identical on every rank (same rows, same deterministic kernels)."""
        s_emb, pi, o_emb, both = self._queries(s, p, o)
        n = s.numel()
        # one stacked launch against the 2n exchanged rows [s ; o]: sp block column n+i is o_i, po block column i is s_i
        x = self.backend.score_sp_po(self.model, s_emb, self.rel, pi, o_emb, both, self.l_norm, self.precision)
        ar = torch.arange(n, device=x.device)
        t_sp = x[ar, n + ar].contiguous()
        t_po = x[ar, 2 * n + ar].contiguous()
        return (t_sp, t_po), (s_emb, pi, o_emb)

    def rank_sp_po(self, s, p, o, filter_sp=None, filter_po=None, rtol=1e-4, atol=1e-5, filter_csr=None):
        """(s_rank, s_ties, o_rank, o_ties) over ALL entities; filter_* are this rank's column slices
        [n, E_local] of the reference's +inf label matrix (eval_entity_ranking.py:287-290,561-566), or — without any
        dense matrix — filter_csr = (offsets [2n+1], columns): the known answers of the 2n stacked rows (sp_ rows, then
        _po rows) as GLOBAL entity ids, sorted per row; this rank keeps the ids of its shard (the row's own answer
        stays in: :287-290).
        One stacked score+rank launch on the shard, then an integer all-reduce => bit-exact, 32*n bytes instead
        of moving logits."""
        (t_sp, t_po), (s_emb, pi, o_emb) = self.true_scores(s, p, o)
        n = s.numel()
        dev = self.ent.device
        counts = torch.zeros((2, 2 * n), dtype=torch.int64, device=dev)
        if self.hi > self.lo:
            filt = None
            if filter_csr is not None:
                off, col = filter_csr
                rows = torch.repeat_interleave(torch.arange(2 * n, device=dev), off[1:] - off[:-1])
                keep = (col >= self.lo) & (col < self.hi)
                loff = torch.zeros(2 * n + 1, dtype=torch.int64, device=dev)
                loff[1:] = torch.cumsum(torch.bincount(rows[keep], minlength=2 * n), 0)
                own = torch.cat([o.long(), s.long()]) - self.lo
                filt = (loff, (col[keep] - self.lo).contiguous(), own.contiguous())
            elif filter_sp is not None or filter_po is not None:
                z = lambda f: f if f is not None else torch.zeros((n, self.hi - self.lo), device=dev)
                filt = torch.cat([z(filter_sp), z(filter_po)], 0)
            r, t = self.backend.rank_sp_po(self.model, s_emb, self.rel, pi, o_emb, self.ent,
                                           torch.cat([t_sp, t_po]), filt, rtol, atol, self.l_norm, self.precision)
            counts[0], counts[1] = r, t
        self._all_reduce(counts)
        # rows 0..n-1 ranked the objects (sp_), rows n..2n-1 the subjects (_po)
        return counts[0, n:], counts[1, n:], counts[0, :n], counts[1, :n]

    # -- duck-typed model interface of kge_b200.evaluate.EntityRankingEvaluator ------------------------
    # (filtered entity ranking over a sharded table: every rank runs the same evaluator on the same batches;
    # use chunk_size=-1 — the shards are the chunks)
    def score_sp(self, s, p, o=None):
        """[n, |o|] scores of (s,p) against the entities `o` (global ids), identical on every rank: the
        true-score path of the evaluation loop (eval_entity_ranking.py:192-203)."""
        if o is None:
            return self.score_sp_po(s, p, s)[:, : self.E]
        rows = self.gather_entity_rows(torch.cat([s.long(), o.long()]))
        n = s.numel()
        return self.backend.score_1vsN(self.model, "sp_", rows[:n], self.rel[p.long()], rows[n:], self.l_norm, self.precision)

    def score_po(self, p, o, s=None):
        if s is None:
            return self.score_sp_po(o, p, o)[:, self.E:]
        rows = self.gather_entity_rows(torch.cat([o.long(), s.long()]))
        n = o.numel()
        return self.backend.score_1vsN(self.model, "_po", rows[:n], self.rel[p.long()], rows[n:], self.l_norm, self.precision)

    def _rank_dir(self, combine, q, p, true_scores, entity_subset, filter_labels, rtol, atol, rank, ties):
        if entity_subset is not None:
            raise ValueError("the sharded model ranks against its whole shard: use chunk_size=-1")
        q_emb = self.gather_entity_rows(q)
        n = q.numel()
        counts = torch.zeros((2, n), dtype=torch.int64, device=self.ent.device)
        if self.hi > self.lo:
            f = None if filter_labels is None else filter_labels[:, self.lo:self.hi].contiguous()
            r, t = self.backend.rank_1vsN(self.model, combine, q_emb, self.rel[p.long()], self.ent, true_scores, f, rtol,
                                          atol, self.l_norm, self.precision)
            counts[0], counts[1] = r, t
        self._all_reduce(counts)
        if rank is None:
            rank = torch.zeros(n, dtype=torch.int64, device=self.ent.device)
        if ties is None:
            ties = torch.zeros(n, dtype=torch.int64, device=self.ent.device)
        rank += counts[0]
        ties += counts[1]
        return rank, ties

    def rank_sp(self, s, p, true_scores, entity_subset=None, filter_labels=None, rtol=1e-4, atol=1e-5, rank=None, ties=None):
        return self._rank_dir("sp_", s, p, true_scores, entity_subset, filter_labels, rtol, atol, rank, ties)

    def rank_po(self, p, o, true_scores, entity_subset=None, filter_labels=None, rtol=1e-4, atol=1e-5, rank=None, ties=None):
        return self._rank_dir("_po", o, p, true_scores, entity_subset, filter_labels, rtol, atol, rank, ties)

    # -- 1vsAll BCE ------------------------------------------------------------------------------------
    def loss_1vsall_bce(self, s, p, o, offset: float = 0.0):
        """(BCE(score_sp, o) + BCE(score_po, s)) / n with sum reductions (train_1vsAll.py:48-82): BCE is
        additive over columns, so each rank reduces its shard and one scalar is all-reduced."""
        s_emb, pi, o_emb, _ = self._queries(s, p, o)
        p_emb = self.rel[pi]
        n = s.numel()

        def local(idx):
            idx = idx.long()
            l = idx - self.lo
            return torch.where((idx >= self.lo) & (idx < self.hi), l, torch.full_like(l, -1))

        tot = torch.zeros((), dtype=torch.float32, device=self.ent.device)
        if self.hi > self.lo:
            tot = tot + self.backend.bce_1vsN(self.model, "sp_", s_emb, p_emb, self.ent, local(o), offset,
                                              self.l_norm, self.precision)
            tot = tot + self.backend.bce_1vsN(self.model, "_po", o_emb, p_emb, self.ent, local(s), offset,
                                              self.l_norm, self.precision)
        return self._all_reduce(tot) / n

    # -- top-k -----------------------------------------------------------------------------------------
    def topk_sp(self, s, p, k: int):
        """Global top-k objects per (s,p): local top-k on the shard's logits, all-gather of the [n,k]
        (value, global index) pairs, merge; ties broken towards the lowest entity id."""
        s_emb = self.gather_entity_rows(s)
        p_emb = self.rel[p.long()]
        n = s.numel()
        dev = self.ent.device
        vals = torch.full((n, k), float("-inf"), dtype=torch.float32, device=dev)
        idxs = torch.full((n, k), self.E, dtype=torch.int64, device=dev)
        m = self.hi - self.lo
        if m > 0:
            loc = self.backend.score_1vsN(self.model, "sp_", s_emb, p_emb, self.ent, self.l_norm, self.precision)
            kk = min(k, m)
            v, i = _topk_lowest_index(loc, kk)
            vals[:, :kk], idxs[:, :kk] = v, i + self.lo
        if self.world > 1:
            vs = [torch.empty_like(vals) for _ in range(self.world)]
            is_ = [torch.empty_like(idxs) for _ in range(self.world)]
            dist.all_gather(vs, vals, group=self.group)
            dist.all_gather(is_, idxs, group=self.group)
            vals, idxs = torch.cat(vs, 1), torch.cat(is_, 1)
        v, pos = _topk_lowest_index(vals, k, idxs)
        return v, torch.gather(idxs, 1, pos)


def _topk_lowest_index(values: torch.Tensor, k: int, ids: Optional[torch.Tensor] = None):
    """top-k by value, ties broken by the lowest id (column position if ids is None): a stable sort
    on -value after ordering columns by id gives a deterministic, shard-count-independent result."""
    if ids is None:
        order = torch.sort(-values, dim=1, stable=True).indices[:, :k]
        return torch.gather(values, 1, order), order
    by_id = torch.sort(ids, dim=1, stable=True).indices
    v = torch.gather(values, 1, by_id)
    order = torch.sort(-v, dim=1, stable=True).indices[:, :k]
    pos = torch.gather(by_id, 1, order)
    return torch.gather(values, 1, pos), pos
