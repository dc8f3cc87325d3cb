"""Model-level GPU tests in the shape of the reference's own tests/test_model.py:29-71
(score_spo vs score_sp vs score_po consistency for every model) plus the mirror classes."""
import pytest
import torch

from oracle import kge_oracle as orc

pytestmark = pytest.mark.gpu

MODELS = ["complex", "distmult", "simple", "cp", "rescal", "transe", "rotate"]


@pytest.mark.parametrize("name", MODELS)
def test_score_equality(name):
    """test_model.py:29-71 — all (s,p,o) on a 4-entity/3-relation graph, dim 32."""
    from kge_b200 import KgeModel

    E, R = 4, 3
    D = 32
    model = KgeModel(name, E, R, D, seed=0).cuda()
    dev = "cuda"
    s = torch.arange(E, device=dev).repeat_interleave(R * E)
    p = torch.arange(R, device=dev).repeat_interleave(E).repeat(E)
    o = torch.arange(E, device=dev).repeat(R * E)
    spo_s = model.score_spo(s, p, o, direction="s")
    spo_o = model.score_spo(s, p, o, direction="o")
    s2 = torch.arange(E, device=dev).repeat_interleave(R)
    p2 = torch.arange(R, device=dev).repeat(E)
    sp = model.score_sp(s2, p2).contiguous()
    assert torch.allclose(spo_o.view(-1), sp.view(-1), atol=1e-5, rtol=1e-4)
    p3 = torch.arange(R, device=dev).repeat_interleave(E)
    o3 = torch.arange(E, device=dev).repeat(R)
    po = model.score_po(p3, o3).t().contiguous()
    assert torch.allclose(spo_s.view(-1), po.view(-1), atol=1e-5, rtol=1e-4)


@pytest.mark.parametrize("name", MODELS)
def test_model_methods_vs_oracle(name):
    from kge_b200 import KgeModel

    E, R, D, n = 517, 6, 64 if name != "rescal" else 24, 37
    model = KgeModel(name, E, R, D, seed=1, sigma=0.5).cuda()
    ent = model._entity_embedder.weight.detach().cpu()
    rel = model._relation_embedder.weight.detach().cpu()
    tri = orc.make_triples(E, R, n)
    t = tri.cuda()
    s, p, o = t[:, 0], t[:, 1], t[:, 2]
    g = torch.Generator().manual_seed(0)
    sub = torch.randperm(E, generator=g)[:100]
    psub = torch.randperm(R, generator=g)[:3]

    def close(got, ref, what):
        rms = max(float(ref.pow(2).mean().sqrt()), 1e-6)
        err = float((got.cpu() - ref).abs().max())
        assert err <= 1e-4 * rms, (name, what, err, rms)

    close(model.score_spo(s, p, o), orc.score_spo(name, ent, rel, tri[:, 0], tri[:, 1], tri[:, 2]), "spo")
    close(model.score_sp(s, p), orc.score_sp(name, ent, rel, tri[:, 0], tri[:, 1]), "sp")
    close(model.score_po(p, o, sub.cuda()), orc.score_po(name, ent, rel, tri[:, 1], tri[:, 2], sub), "po subset")
    close(model.score_so(s, o), orc.score_so(name, ent, rel, tri[:, 0], tri[:, 2]), "so")
    close(model.score_so(s, o, psub.cuda()), orc.score_so(name, ent, rel, tri[:, 0], tri[:, 2], psub), "so subset")
    close(model.score_sp_po(s, p, o, sub.cuda()), orc.score_sp_po(name, ent, rel, tri[:, 0], tri[:, 1], tri[:, 2], sub), "sp_po subset")
    sc = model.get_scorer()
    close(sc.score_emb(ent.cuda()[s], rel.cuda()[p], ent.cuda()[sub.cuda()], "sp_"),
          orc.score_sp(name, ent, rel, tri[:, 0], tri[:, 1], sub), "score_emb sp_")
    close(sc.score_emb(ent.cuda()[s], rel.cuda()[psub.cuda()], ent.cuda()[o], "s_o"),
          orc.score_so(name, ent, rel, tri[:, 0], tri[:, 2], psub), "score_emb s_o")
    with pytest.raises(ValueError):
        sc.score_emb(ent.cuda()[s], rel.cuda()[p], ent.cuda()[o], "xyz")


def test_loss_ranking_sampler_mirrors():
    from kge_b200 import BatchNegativeSample, KgeLoss, KgeModel
    from kge_b200 import model as km

    name, E, R, D, n, K = "complex", 700, 5, 64, 50, 33
    m = KgeModel(name, E, R, D, seed=2, sigma=0.5).cuda()
    ent = m._entity_embedder.weight.detach().cpu()
    rel = m._relation_embedder.weight.detach().cpu()
    tri = orc.make_triples(E, R, n)
    t = tri.cuda()
    scores = m.score_sp(t[:, 0], t[:, 1])
    for kind, fn in (("bce", orc.bce_loss), ("kl", orc.kl_loss)):
        got = float(KgeLoss.create(kind)(scores, t[:, 2]))
        ref = float(fn(scores.cpu(), tri[:, 2]))
        assert abs(got - ref) <= 1e-5 * abs(ref)
        fused = float(m.score_sp_loss(t[:, 0], t[:, 1], t[:, 2], kind))
        assert abs(fused - ref) <= 1e-4 * abs(ref)
    with pytest.raises(ValueError):
        KgeLoss.create("margin_ranking")
    g = torch.Generator().manual_seed(0)
    neg = torch.randint(0, E, (n, K), generator=g)
    bns = BatchNegativeSample(tri, 2, neg).to("cuda")
    ref = orc.ns_score(name, ent, rel, tri, neg, 2, "batch")
    got = bns.score(m)
    assert float((got.cpu() - ref).abs().max()) <= 1e-4 * float(ref.pow(2).mean().sqrt())
    full = bns.score_with_positive(m)
    assert full.shape == (n, K + 1)
    # ranking arithmetic mirrors
    true = scores[torch.arange(n, device="cuda"), t[:, 2]]
    r, ti = km.get_ranks_and_num_ties(scores, true)
    rr, tt = orc.ranks_and_ties(scores.cpu(), true.cpu())
    assert torch.equal(r.cpu(), rr) and torch.equal(ti.cpu(), tt)
    assert torch.equal(km.get_ranks(r, ti).cpu(), orc.final_ranks(rr, tt))
    fr, ft = m.rank_sp(t[:, 0], t[:, 1], true)
    assert torch.equal(fr.cpu(), rr) and torch.equal(ft.cpu(), tt)
