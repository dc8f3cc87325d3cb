"""Host logic of the entity-sharded path with world_size=2 on CPU (gloo): ownership, query-row
exchange, logits all-gather, integer rank all-reduce, BCE all-reduce, top-k merge.  The local scorer
is the CPU oracle (injected backend): the CUDA backend is exercised by tests/test_gpu_sharded.py."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import kge_oracle as orc


class OracleBackend:
    def exchange_rows(self, shard, lo, idx):
        out = torch.zeros((idx.numel(), shard.shape[1]))
        mine = (idx >= lo) & (idx < lo + shard.shape[0])
        out[mine] = shard[idx[mine] - lo]
        return out

    def score_sp_po(self, model, s_emb, rel, p, o_emb, cand, l_norm, precision, out=None):
        x = torch.cat([orc.score_emb(model, s_emb, rel[p], cand, "sp_", l_norm),
                       orc.score_emb(model, cand, rel[p], o_emb, "_po", l_norm)], 1)
        if out is not None:
            out[:, : x.shape[1]] = x
            return out
        return x

    def rank_sp_po(self, model, s_emb, rel, p, o_emb, cand, true2n, filter2n, rtol, atol, l_norm, precision):
        n = s_emb.shape[0]
        if isinstance(filter2n, tuple):          # CSR -> the dense matrix the kernels never build
            off, col, own = filter2n
            dense = torch.zeros((2 * n, cand.shape[0]))
            rows = torch.repeat_interleave(torch.arange(2 * n), off[1:] - off[:-1])
            dense[rows, col] = float("inf")
            ok = (own >= 0) & (own < cand.shape[0])
            dense[torch.arange(2 * n)[ok], own[ok]] = 0.0
            filter2n = dense
        x = torch.cat([orc.score_emb(model, s_emb, rel[p], cand, "sp_", l_norm),
                       orc.score_emb(model, cand, rel[p], o_emb, "_po", l_norm)], 0)
        if filter2n is not None:
            x = x - filter2n
        return orc.ranks_and_ties(x, true2n, rtol, atol)

    def score_1vsN(self, model, combine, q, p, cand, l_norm, precision):
        if combine == "sp_":
            return orc.score_emb(model, q, p, cand, "sp_", l_norm)
        return orc.score_emb(model, cand, p, q, "_po", l_norm)

    def score_spo(self, model, s, p, o, l_norm):
        return orc.score_emb(model, s, p, o, "spo", l_norm).view(-1)

    def rank_1vsN(self, model, combine, q, p, cand, true, filt, rtol, atol, l_norm, precision):
        x = self.score_1vsN(model, combine, q, p, cand, l_norm, precision)
        if filt is not None:
            x = x - filt
        return orc.ranks_and_ties(x, true, rtol, atol)

    def bce_1vsN(self, model, combine, q, p, cand, local_labels, offset, l_norm, precision):
        x = self.score_1vsN(model, combine, q, p, cand, l_norm, precision) + offset
        tot = torch.nn.functional.softplus(x).sum()
        has = local_labels >= 0
        rows = torch.arange(x.shape[0])[has]
        return tot - x[rows, local_labels[has]].sum()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, model, E, R, D, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from kge_b200.sharded import ShardedKgeModel

        torch.set_num_threads(2)
        ent, rel = orc.make_tables(model, E, R, D, sigma=0.7)
        tri = orc.make_triples(E, R, n, seed=5)
        s, p, o = tri[:, 0], tri[:, 1], tri[:, 2]
        lo, hi = ShardedKgeModel.shard_bounds(E, world, rank)
        m = ShardedKgeModel(model, ent[lo:hi].clone(), rel, E, rank, world, backend=OracleBackend())
        res = {}
        # 1. query rows are exchanged exactly
        assert torch.equal(m.gather_entity_rows(s), ent[s])
        # 2. logits == single-process oracle
        full = m.score_sp_po(s, p, o)
        ref = orc.score_sp_po(model, ent, rel, s, p, o)
        assert full.shape == ref.shape
        res["logit_err"] = float((full - ref).abs().max() / ref.abs().max())
        # 3. ranks: additive integer counts == ranks of the gathered logits
        g = torch.Generator().manual_seed(11)
        filt = torch.zeros((n, 2 * E))
        filt[torch.rand((n, 2 * E), generator=g) < 0.03] = float("inf")
        filt[torch.arange(n), o] = 0.0
        filt[torch.arange(n), E + s] = 0.0
        (t_sp, t_po), _ = m.true_scores(s, p, o)
        # the true scores come from the 1-vs-N path: they ARE entries of the gathered logits (ties >= 1 by construction)
        ar = torch.arange(n)
        assert torch.equal(t_sp, full[ar, o]) and torch.equal(t_po, full[ar, E + s])
        for f in (None, filt):
            fsp = None if f is None else f[:, lo:hi].contiguous()
            fpo = None if f is None else f[:, E + lo:E + hi].contiguous()
            s_rank, s_ties, o_rank, o_ties = m.rank_sp_po(s, p, o, fsp, fpo)
            sp, po = full[:, :E], full[:, E:]
            if f is not None:
                sp, po = sp - f[:, :E], po - f[:, E:]
            rr, tt = orc.ranks_and_ties(sp, t_sp)
            assert torch.equal(o_rank, rr) and torch.equal(o_ties, tt)
            rr, tt = orc.ranks_and_ties(po, t_po)
            assert torch.equal(s_rank, rr) and torch.equal(s_ties, tt)
            assert int(o_ties.min()) >= 1 and int(s_ties.min()) >= 1
            if f is not None:
                # the same filter as CSR over the stacked rows (global ids; the own answers included, as the index has them)
                fs = torch.cat([f[:, :E], f[:, E:]], 0).clone()
                fs[torch.arange(n), o] = float("inf")
                fs[n + torch.arange(n), s] = float("inf")
                rr_, cc_ = torch.nonzero(torch.isinf(fs), as_tuple=True)
                offc = torch.zeros(2 * n + 1, dtype=torch.int64)
                offc[1:] = torch.cumsum(torch.bincount(rr_, minlength=2 * n), 0)
                got = m.rank_sp_po(s, p, o, filter_csr=(offc, cc_))
                for x, y in zip(got, (s_rank, s_ties, o_rank, o_ties)):
                    assert torch.equal(x, y)
        # 4. BCE
        got = float(m.loss_1vsall_bce(s, p, o, 0.5))
        want = float(orc.train_1vsall_forward(model, ent, rel, tri, "bce", 0.5))
        res["bce_rel"] = abs(got - want) / abs(want)
        # 5. top-k with deterministic tie-break (duplicate rows => exact score ties across shards)
        v, i = m.topk_sp(s, p, 7)
        refsp = orc.score_sp(model, ent, rel, s, p)
        order = torch.sort(-refsp, dim=1, stable=True).indices[:, :7]
        res["topk_idx_match"] = float((i == order).float().mean())
        res["topk_val_err"] = float((v - torch.gather(refsp, 1, order)).abs().max())
        if rank == 0:
            out.put(res)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("model,D", [("complex", 32), ("transe", 32), ("rescal", 12)])
def test_sharded_world2_gloo(model, D):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    E, R, n = 203, 5, 17          # odd E: shards of 102 and 101 rows
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, model, E, R, D, n, out)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(120)
        assert pr.exitcode == 0, f"worker failed with exit code {pr.exitcode}"
    res = out.get()
    assert res["logit_err"] < 1e-5, res
    assert res["bce_rel"] < 1e-5, res
    assert res["topk_idx_match"] > 0.99 and res["topk_val_err"] < 1e-4, res


def test_shard_bounds_cover_everything():
    from kge_b200.sharded import ShardedKgeModel

    for E, G in ((10, 3), (4800000, 8), (7, 8), (203, 2)):
        spans = [ShardedKgeModel.shard_bounds(E, G, r) for r in range(G)]
        assert spans[0][0] == 0 and spans[-1][1] == E
        for a, b in zip(spans, spans[1:]):
            assert a[1] == b[0]


def _eval_worker(rank, world, port, model, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import numpy as np
        from kge_b200.evaluate import EntityRankingEvaluator
        from kge_b200.sharded import ShardedKgeModel

        torch.set_num_threads(2)
        z = np.load(os.path.join(os.path.dirname(__file__), "golden", f"jobs_{model}.npz"))
        g = {k: torch.from_numpy(z[k]) if z[k].ndim else z[k].item() for k in z.files}
        E = g["ent"].shape[0]
        lo, hi = ShardedKgeModel.shard_bounds(E, world, rank)
        m = ShardedKgeModel(model, g["ent"][lo:hi].clone(), g["rel"], E, rank, world, backend=OracleBackend())
        ev = EntityRankingEvaluator(m, E, [g["train"], g["valid"]], g["test"], batch_size=16, hits_at_k_s=(1, 3, 10))
        met = ev.evaluate(g["valid"])
        if rank == 0:
            out.put((met, {k: float(v) for k, v in g.items() if k.startswith("valid_")}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("model", ["complex", "transe"])
def test_sharded_evaluation_matches_reference_job(model):
    """The evaluation loop over an entity-sharded table (2 ranks, gloo): integer rank counts are all-reduced per
    (ranking, direction); metrics equal the reference EntityRankingJob's trace."""
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_eval_worker, args=(r, world, port, model, out)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(180)
        assert pr.exitcode == 0, f"worker failed with exit code {pr.exitcode}"
    met, want = out.get()
    for k, v in want.items():
        assert met[k[len("valid_"):]] == pytest.approx(v, rel=1e-6, abs=1e-9), (k, met[k[len("valid_"):]], v)
