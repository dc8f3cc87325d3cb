"""BASELINE.json configurations at their full shapes (parity-test cases, not bench lines):
direct oracle comparison on a bounded sample of query rows, plus size-independent properties
(spo == gathered sp_/_po entries as in the reference's tests/test_model.py:29-71; linearity of the
dot family; additivity of rank counts over entity chunks; score_sp_po == [score_sp | score_po])."""
import pytest
import torch

from oracle import kge_oracle as orc

pytestmark = pytest.mark.gpu
S, P, O = 0, 1, 2


def _close(got, ref, what, tol=1e-4):
    got, ref = got.detach().cpu().double(), ref.double()
    rms = max(float(ref.pow(2).mean().sqrt()), 1e-6)
    err = float((got - ref).abs().max())
    assert err <= tol * rms, f"{what}: max|d|={err:.3e} rms={rms:.3e}"


@pytest.fixture(scope="module")
def eng():
    from kge_b200 import engine
    return engine


def test_cfg2_complex_fb15k237_shape(eng):
    """ComplEx d=512 1vsAll+BCE, 14 541 entities / 237 relations, n=1024 (the bench workload)."""
    model, E, R, D, n = "complex", 14541, 237, 512, 1024
    ent, rel = orc.make_tables(model, E, R, D)
    tri = orc.make_triples(E, R, n)
    ce, cr, ct = ent.cuda(), rel.cuda(), tri.cuda()
    s, p, o = ct[:, S].contiguous(), ct[:, P].contiguous(), ct[:, O].contiguous()
    full = eng.score_sp_po(model, ce, cr, s, p, o)
    _close(full, orc.score_sp_po(model, ent, rel, tri[:, S], tri[:, P], tri[:, O]), "sp_po")
    # [sp | po] halves equal the separate calls bit-for-bit (same kernel, stacked rows)
    assert torch.equal(full[:, :E], eng.score_1vsN(model, "sp_", ce, cr, ce, s, p))
    assert torch.equal(full[:, E:], eng.score_1vsN(model, "_po", ce, cr, ce, o, p))
    # spo consistency (reference test_score_equality tolerances)
    spo = eng.score_spo(model, ce, cr, ce, s, p, o)
    idx = torch.arange(n, device="cuda")
    assert torch.allclose(spo, full[idx, o], atol=1e-3, rtol=1e-4)
    assert torch.allclose(spo, full[idx, E + s], atol=1e-3, rtol=1e-4)
    # fused step loss for both losses
    for loss in ("bce", "kl"):
        ref = float(orc.train_1vsall_forward(model, ent, rel, tri, loss))
        got = float(eng.train_1vsall_forward(model, ce, cr, ct, loss))
        assert abs(got - ref) <= 1e-4 * abs(ref), (loss, got, ref)
    # linearity of the dot family in the entity table: score(2T) == 2 * score(T) exactly (power of two)
    twice = eng.score_1vsN(model, "sp_", ce, cr, (2.0 * ce), s, p)
    assert torch.equal(twice, 2.0 * full[:, :E])
    # rank counts are additive over entity chunks (eval_entity_ranking.py:222-229,310-313)
    true = full[idx, o].clone()
    r_all, t_all = eng.score_1vsN_rank(model, "sp_", ce, cr, ce, true, s, p)
    r = torch.zeros(n, dtype=torch.int64, device="cuda")
    t = torch.zeros(n, dtype=torch.int64, device="cuda")
    for c0 in range(0, E, 5000):
        sub = torch.arange(c0, min(c0 + 5000, E), device="cuda")
        eng.score_1vsN_rank(model, "sp_", ce, cr, ce, true, s, p, sub, rank=r, ties=t)
    assert int((r - r_all).abs().max()) <= 1 and int((t - t_all).abs().max()) <= 1   # tolerance-band flips only
    assert int(t_all.min()) >= 1          # the true answer always ties with itself


def test_cfg3_rotate_ns_wn18rr_shape(eng):
    """RotatE d=512, negative sampling K=1000 (s and o slots), 40 943 entities."""
    model, E, R, D, n, K = "rotate", 40943, 11, 512, 96, 1000
    ent, rel = orc.make_tables(model, E, R, D)
    tri = orc.make_triples(E, R, n)
    g = torch.Generator().manual_seed(1)
    for slot in (S, O):
        neg = torch.randint(0, E, (n, K), generator=g)
        got = eng.ns_score(model, ent.cuda(), rel.cuda(), tri.cuda(), neg.cuda(), slot, True)
        ref = orc.ns_scores_with_positive(model, ent, rel, tri, neg, slot, "triple")
        _close(got, ref, f"rotate NS slot {slot}")
        # BCE with offset over the [n, 1+K] block, labels [1,0,...]   train_negative_sampling.py:128-156
        lab = orc.ns_labels(n, K)
        gl = float(eng.loss_dense(got, lab.cuda(), "bce", 5.0))
        rl = float(orc.bce_loss(ref, lab, 5.0))
        assert abs(gl - rl) <= 1e-4 * abs(rl)


def test_cfg4_rescal_kvsall_yago_shape(eng):
    """RESCAL d=200 (relation rows of 40 000 floats), KvsAll multi-hot labels, 123 182 entities."""
    model, E, R, D, n = "rescal", 123182, 37, 200, 64
    ent, rel = orc.make_tables(model, E, R, D, sigma=0.3)
    tri = orc.make_triples(E, R, n)
    ce, cr, ct = ent.cuda(), rel.cuda(), tri.cuda()
    s, p, o = ct[:, S].contiguous(), ct[:, P].contiguous(), ct[:, O].contiguous()
    ref = orc.score_sp(model, ent, rel, tri[:, S], tri[:, P])
    _close(eng.score_1vsN(model, "sp_", ce, cr, ce, s, p), ref, "rescal sp")
    _close(eng.score_1vsN(model, "_po", ce, cr, ce, o, p), orc.score_po(model, ent, rel, tri[:, P], tri[:, O]), "rescal po")
    g = torch.Generator().manual_seed(2)
    lab = (torch.rand((n, E), generator=g) < 2e-4).float()
    lab[torch.arange(n), tri[:, O]] = 1.0
    lab = orc.kvsall_smooth_labels(lab, 0.1)
    for loss, fn in (("bce", orc.bce_loss), ("kl", orc.kl_loss)):
        got = float(eng.score_1vsN_loss(model, "sp_", ce, cr, ce, lab.cuda(), s, p, None, loss))
        want = float(fn(ref, lab))
        assert abs(got - want) <= 1e-4 * abs(want), (loss, got, want)


def test_cfg5_transe_wikidata_shard_shape(eng):
    """TransE d=512 L1 against one Wikidata5M-shaped shard (600 000 rows) + rank counting."""
    model, E, R, D, n = "transe", 600000, 822, 512, 12
    ent, rel = orc.make_tables(model, E, R, D)
    tri = orc.make_triples(E, R, n)
    ce, cr, ct = ent.cuda(), rel.cuda(), tri.cuda()
    s, p, o = ct[:, S].contiguous(), ct[:, P].contiguous(), ct[:, O].contiguous()
    ref = orc.score_sp(model, ent, rel, tri[:, S], tri[:, P])
    got = eng.score_1vsN(model, "sp_", ce, cr, ce, s, p)
    _close(got, ref, "transe shard sp")
    true = got[torch.arange(n, device="cuda"), o].clone()
    rr, tt = orc.ranks_and_ties(got.cpu(), true.cpu())
    r, t = eng.score_1vsN_rank(model, "sp_", ce, cr, ce, true, s, p)
    assert torch.equal(r.cpu(), rr) and torch.equal(t.cpu(), tt)


def test_complex_wikidata_shard_shape_tensor_core(eng):
    """ComplEx d=512 against a 600 000-row shard through the tensor-core kernel (2 344 entity tiles)."""
    model, E, R, D, n = "complex", 600000, 822, 512, 160
    ent, rel = orc.make_tables(model, E, R, D)
    tri = orc.make_triples(E, R, n)
    ce, cr, ct = ent.cuda(), rel.cuda(), tri.cuda()
    s, p = ct[:, S].contiguous(), ct[:, P].contiguous()
    _close(eng.score_1vsN(model, "sp_", ce, cr, ce, s, p), orc.score_sp(model, ent, rel, tri[:, S], tri[:, P]),
           "complex shard sp")
    got = float(eng.score_1vsN_loss(model, "sp_", ce, cr, ce, ct[:, O].contiguous(), s, p, None, "kl"))
    want = float(orc.kl_loss(orc.score_sp(model, ent, rel, tri[:, S], tri[:, P]), tri[:, O]))
    assert abs(got - want) <= 1e-4 * abs(want)
