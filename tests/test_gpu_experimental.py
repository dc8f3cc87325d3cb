"""Parity tests for EXPERIMENTAL kernels that are not on the default path yet (selected through
B200KGE_TC_VERSION).  They run only with B200KGE_EXPERIMENTAL=1 so that the regular `-m gpu` suite covers
exactly what ships; the bar is the same (floating point <= 1e-4 * rms, rank/tie counts bit-exact on the
kernel's own scores).

    B200KGE_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental.py -x -q
"""
import os

import numpy as np
import pytest
import torch

from oracle import kge_oracle as orc

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("B200KGE_EXPERIMENTAL") != "1",
                                 reason="experimental kernels: set B200KGE_EXPERIMENTAL=1")]

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
S, P, O = 0, 1, 2
TOL = 1e-4


@pytest.fixture(scope="module")
def eng():
    from kge_b200 import engine

    assert torch.cuda.is_available() and engine.device_ok()
    return engine


VARIANTS = [("3", "0", "64"), ("3", "0", "32"), ("4", "0", "64"), ("4", "1", "64")]
# (B200KGE_TC_VERSION, B200KGE_TC4_DIRECT, B200KGE_TC3_TK)


@pytest.fixture(params=VARIANTS, ids=["tc3", "tc3-tk32", "tc4-forward", "tc4-direct"])
def variant(request, monkeypatch):
    """Selects the experimental kernel for the duration of a test (the default path is restored afterwards)."""
    ver, direct, tk = request.param

    def select():
        monkeypatch.setenv("B200KGE_TC_VERSION", ver)
        monkeypatch.setenv("B200KGE_TC4_DIRECT", direct)
        monkeypatch.setenv("B200KGE_TC3_TK", tk)
    return select


def _load(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(z[k]) if z[k].ndim else z[k].item() for k in z.files}


def _assert_close(got, ref, what, tol=TOL):
    got = got.detach().cpu().double()
    ref = ref.double()
    assert got.shape == ref.shape, f"{what}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    rms = max(float(ref.pow(2).mean().sqrt()), 1e-6)
    err = float((got - ref).abs().max()) if ref.numel() else 0.0
    assert err <= tol * rms, f"{what}: max|d|={err:.3e} rms={rms:.3e} ratio={err / rms:.2e}"


def test_presplit_fp16_golden(eng, variant):
    variant()
    for fname, model in (("scores_complex.npz", "complex"), ("scores_distmult.npz", "distmult"),
                         ("scores_simple.npz", "simple"), ("scores_complex_sigma01.npz", "complex")):
        g = _load(fname)
        ent, rel, tri = g["ent"].cuda(), g["rel"].cuda(), g["triples"].cuda()
        s, p, o = tri[:, S].contiguous(), tri[:, P].contiguous(), tri[:, O].contiguous()
        sub = g["subset"].cuda()
        for prec in ("auto", "tf32+bf16x2"):
            _assert_close(eng.score_1vsN(model, "sp_", ent, rel, ent, s, p, None, precision=prec), g["sp"], fname + " sp")
            _assert_close(eng.score_1vsN(model, "_po", ent, rel, ent, o, p, sub, precision=prec), g["po_subset"], fname + " po_subset")
            _assert_close(eng.score_sp_po(model, ent, rel, s, p, o, None, precision=prec), g["sp_po"], fname + " sp_po")


@pytest.mark.parametrize("sigma", [1.0, 1e-3])
def test_presplit_fp16_medium(eng, variant, sigma):
    """Dense scores, gathered candidate subsets, fused BCE/KL, fused rank counting at ragged sizes (tiles cut in
    both dimensions, K not a multiple of the 64-wide chunk for RESCAL/CP), including tiny-valued tables that a
    fixed fp16 scale would flush."""
    variant()
    for model, D in (("complex", 192), ("distmult", 64), ("simple", 128), ("rescal", 40), ("cp", 200)):
        E, R, n = 6007, 7, 389
        ent, rel = orc.make_tables(model, E, R, D, sigma=sigma)
        tri = orc.make_triples(E, R, n)
        ce, cr, ct = ent.cuda(), rel.cuda(), tri.cuda()
        s, p, o = ct[:, S].contiguous(), ct[:, P].contiguous(), ct[:, O].contiguous()
        ref = orc.score_sp_po(model, ent, rel, tri[:, S], tri[:, P], tri[:, O])
        _assert_close(eng.score_sp_po(model, ce, cr, s, p, o), ref, f"{model} sp_po")
        sub = torch.randperm(E, generator=torch.Generator().manual_seed(1))[:1500]
        got = eng.score_1vsN(model, "_po", ce, cr, ce, o, p, sub.cuda())
        _assert_close(got, orc.score_po(model, ent, rel, tri[:, P], tri[:, O], sub), f"{model} po subset")
        if model == "cp":
            continue        # stacked fused epilogues are not offered for CP
        for loss in ("bce", "kl"):
            refl = float(orc.train_1vsall_forward(model, ent, rel, tri, loss))
            gotl = float(eng.train_1vsall_forward(model, ce, cr, ct, loss))
            assert abs(gotl - refl) <= 1e-4 * abs(refl), (model, loss, gotl, refl)
        dense = eng.score_1vsN(model, "sp_", ce, cr, ce, s, p)
        true = dense[torch.arange(n, device="cuda"), o].clone()
        rr, tt = orc.ranks_and_ties(dense.cpu(), true.cpu())
        r, t = eng.score_1vsN_rank(model, "sp_", ce, cr, ce, true, s, p)
        assert torch.equal(r.cpu(), rr) and torch.equal(t.cpu(), tt)


def test_presplit_fp16_headline_shape(eng, variant):
    """BASELINE configs[1] shape: loss of the experimental path == loss of the default path to 1e-5."""
    E, R, D, n = 14541, 237, 512, 1024
    ent, rel = orc.make_tables("complex", E, R, D)
    tri = orc.make_triples(E, R, n)
    ce, cr, ct = ent.cuda(), rel.cuda(), tri.cuda()
    base = float(eng.train_1vsall_forward("complex", ce, cr, ct, "bce"))
    variant()
    got = float(eng.train_1vsall_forward("complex", ce, cr, ct, "bce"))
    assert abs(got - base) <= 1e-5 * abs(base), (got, base)
    ref = orc.score_sp("complex", ent, rel, tri[:64, S], tri[:64, P])
    _assert_close(eng.score_1vsN("complex", "sp_", ce, cr, ce, ct[:64, S].contiguous(), ct[:64, P].contiguous()), ref,
                  "headline sp")


@pytest.mark.parametrize("model", ["complex", "transe"])
def test_evaluator_on_gpu_matches_reference_job(model):
    """kge_b200.evaluate.EntityRankingEvaluator driving the fused rank kernels reproduces the reference
    EntityRankingJob's trace (host logic is covered on CPU by tests/test_evaluate_cpu.py; this adds the device
    side: chunked subsets, dense filter planes, accumulation into rank/ties).  To be promoted into
    tests/test_gpu_model.py once it has run green on a B200."""
    from kge_b200 import KgeModel
    from kge_b200.evaluate import EntityRankingEvaluator

    g = _load(f"jobs_{model}.npz")
    E, D = g["ent"].shape
    m = KgeModel(model, E, g["rel"].shape[0], D).cuda()
    with torch.no_grad():
        m.get_s_embedder().weight.copy_(g["ent"].cuda())
        m.get_p_embedder().weight.copy_(g["rel"].cuda())
    for bs, chunk in ((16, -1), (100, 7)):
        ev = EntityRankingEvaluator(m, E, [g["train"], g["valid"]], g["test"], batch_size=bs, chunk_size=chunk,
                                    hits_at_k_s=(1, 3, 10), device="cuda")
        met = ev.evaluate(g["valid"])
        for suffix in ("", "_filtered", "_filtered_with_test"):
            for k in ("mean_rank", "mean_reciprocal_rank", "hits_at_1", "hits_at_3", "hits_at_10"):
                want = float(g["valid_" + k + suffix])
                assert abs(met[k + suffix] - want) <= 1e-6 * max(1.0, abs(want)), (k + suffix, met[k + suffix], want)


def test_x_gemm_nt_vs_fp64(eng):
    """Pre-split fp16 GEMM (the backward's building block; also exercises presplit + pairwise_tc3 on shapes
    the scorer never sees: long reductions, few rows, K not a multiple of 64, tiny and huge magnitudes)."""
    g = torch.Generator().manual_seed(0)
    for M, N, K, sa, sb in ((300, 500, 1000, 1.0, 1.0), (2048, 512, 14541, 1e-3, 1.0), (130, 40, 72, 50.0, 1e-4),
                            (5000, 384, 2048, 1.0, 1.0)):
        a = torch.randn((M, K), generator=g) * sa
        b = torch.randn((N, K), generator=g) * sb
        ref = a.double() @ b.double().t()
        got = eng.x_gemm_nt(a.cuda(), b.cuda())
        _assert_close(got, ref, f"gemm {M}x{N}x{K}")


@pytest.mark.parametrize("fname", ["grads_complex_bce.npz", "grads_distmult_bce.npz", "grads_simple_bce.npz",
                                   "grads_cp_bce.npz", "grads_rescal_bce.npz", "grads_complex_kl.npz",
                                   "grads_rescal_kl.npz"])
def test_x_backward_golden(eng, fname):
    """Table gradients of one 1vsAll step (BCE with offset, KL) against the live reference's backward."""
    g = _load(fname)
    model, loss = fname[len("grads_"):-4].split("_")
    d_ent, d_rel = eng.x_train_1vsall_backward(model, g["ent"].cuda(), g["rel"].cuda(), g["triples"].cuda(), loss,
                                               float(g["offset"]))
    _assert_close(d_ent, g["d_ent"], fname + " d_ent")
    _assert_close(d_rel, g["d_rel"], fname + " d_rel")


@pytest.mark.parametrize("loss", ["bce", "kl"])
@pytest.mark.parametrize("model,D", [("complex", 128), ("distmult", 64), ("simple", 128), ("cp", 64), ("rescal", 24)])
def test_x_backward_medium(eng, model, D, loss):
    """Ragged medium shapes with duplicate rows, against the analytic CPU assembly (oracle/kge_fold.py, itself
    pinned to autograd and to the reference's gradients)."""
    from oracle import kge_fold as kf

    E, R, n = 3001, 7, 333
    ent, rel = orc.make_tables(model, E, R, D, sigma=0.5)
    tri = orc.make_triples(E, R, n)
    tri[5] = tri[4]
    off = 0.5 if loss == "bce" else 0.0
    ref_e, ref_r = kf.train_1vsall_backward(model, ent.double(), rel.double(), tri, loss, off)
    d_ent, d_rel = eng.x_train_1vsall_backward(model, ent.cuda(), rel.cuda(), tri.cuda(), loss, off)
    _assert_close(d_ent, ref_e, f"{model} d_ent")
    _assert_close(d_rel, ref_r, f"{model} d_rel")


@pytest.mark.parametrize("model", ["complex", "transe"])
def test_job_traces_on_gpu(eng, model):
    """Job-level traces of the reference (tests/golden/jobs_*.npz) through validated entry points only: 1vsAll
    epoch loss, KvsAll epochs with multi-hot / smoothed labels.  Gated until it has run once on a B200."""
    g = _load(f"jobs_{model}.npz")
    ent, rel, train = g["ent"].cuda(), g["rel"].cuda(), g["train"].long().cuda()
    E = ent.shape[0]
    for loss in ("bce", "kl"):
        got = float(eng.train_1vsall_forward(model, ent, rel, train, loss))
        want = float(g[f"avg_loss_{loss}"])
        assert abs(got - want) <= 1e-4 * abs(want), (loss, got, want)

    def examples(key_cols, val_col):
        keys, inv = torch.unique(train[:, key_cols], dim=0, return_inverse=True)
        labels = torch.zeros((keys.shape[0], E), device="cuda")
        labels.index_put_((inv, train[:, val_col]), torch.ones(len(train), device="cuda"), accumulate=True)
        return keys, labels

    sp_keys, sp_lab = examples([S, P], O)
    po_keys, po_lab = examples([P, O], S)
    n = sp_keys.shape[0] + po_keys.shape[0]
    for loss, eps in (("kl", 0.0), ("kl", 0.2), ("bce", 0.2)):
        lab = (lambda y: (1.0 - eps) * y + 1.0 / E) if eps > 0 else (lambda y: y)
        l_sp = eng.score_1vsN_loss(model, "sp_", ent, rel, ent, lab(sp_lab), sp_keys[:, 0].contiguous(),
                                   sp_keys[:, 1].contiguous(), None, loss, 0.0)
        l_po = eng.score_1vsN_loss(model, "_po", ent, rel, ent, lab(po_lab), po_keys[:, 1].contiguous(),
                                   po_keys[:, 0].contiguous(), None, loss, 0.0)
        got = (float(l_sp) + float(l_po)) / n
        want = float(g[f"kvsall_avg_loss_{loss}_{int(eps * 10)}"])
        assert abs(got - want) <= 1e-4 * abs(want), (loss, eps, got, want)


@pytest.mark.parametrize("model", ["complex", "rotate"])
def test_ns_job_batch_on_gpu(eng, model):
    g = _load(f"nsjob_{model}.npz")
    ent, rel, tri = g["ent"].cuda(), g["rel"].cuda(), g["triples"].long().cuda()
    n, off = tri.shape[0], float(g["offset"])
    total = 0.0
    for slot, nm in ((S, "s"), (P, "p"), (O, "o")):
        neg = g[f"neg_{nm}"].long().cuda()
        scores = eng.ns_score(model, ent, rel, tri, neg, slot, True)
        labels = torch.zeros_like(scores)
        labels[:, 0] = 1.0
        total += float(eng.loss_dense(scores, labels, "bce", off)) / n
    assert abs(total - float(g["avg_loss"])) <= 1e-4 * abs(float(g["avg_loss"])), (total, float(g["avg_loss"]))


@pytest.mark.parametrize("base", ["complex", "transe"])
def test_reciprocal_model_on_gpu(base):
    """kge_b200.ReciprocalRelationsModel (index arithmetic over validated `sp_` entry points) against the live
    reference's ReciprocalRelationsModel.  Gated until it has run once on a B200."""
    from kge_b200 import ReciprocalRelationsModel

    g = _load(f"reciprocal_{base}.npz")
    E, D = g["ent"].shape
    R = int(g["num_relations"])
    m = ReciprocalRelationsModel(base, E, R, D).cuda()
    with torch.no_grad():
        m.get_s_embedder().weight.copy_(g["ent"].cuda())
        m.get_p_embedder().weight.copy_(g["rel2"].cuda())
    tri, sub = g["triples"].long().cuda(), g["subset"].long().cuda()
    s, p, o = tri[:, S].contiguous(), tri[:, P].contiguous(), tri[:, O].contiguous()
    _assert_close(m.score_spo(s, p, o, "o"), g["spo_o"], "spo o")
    _assert_close(m.score_spo(s, p, o, "s"), g["spo_s"], "spo s")
    _assert_close(m.score_sp(s, p), g["sp"], "sp")
    _assert_close(m.score_po(p, o), g["po"], "po")
    _assert_close(m.score_po(p, o, sub), g["po_subset"], "po subset")
    _assert_close(m.score_sp_po(s, p, o), g["sp_po"], "sp_po")
    _assert_close(m.score_sp_po(s, p, o, sub), g["sp_po_subset"], "sp_po subset")


def test_x_penalties_and_normalisation_golden(eng):
    """Row kernels for Lp / N3 penalties and normalisation against the live reference (penalties.npz)."""
    import sys
    sys.path.insert(0, GOLDEN)
    from gen_golden import PENALTY_CASES

    g = _load("penalties.npz")
    for tag, model, eo, ro in PENALTY_CASES:
        space = "complex" if model == "complex" else "euclidean"
        ent, rel, tri = g[f"{tag}_ent"].cuda(), g[f"{tag}_rel"].cuda(), g[f"{tag}_triples"].long().cuda()

        def pen(w, o, idx):
            return float(eng.x_lookup_penalty(w, o["regularize"], o["regularize_weight"], float(o["p"]), o["weighted"],
                                              idx if o["weighted"] else None, space))

        total = pen(rel, ro, tri[:, P])
        total += pen(ent, eo, tri[:, [S, O]]) if eo["weighted"] else 2.0 * pen(ent, eo, None)
        want = float(g[f"{tag}_total"])
        assert abs(total - want) <= 1e-5 * abs(want), (tag, total, want)
    for pn in (1, 2):
        w = g["normalize_in"].cuda().clone()
        eng.x_normalize_rows_(w, float(pn))
        assert torch.allclose(w.cpu(), g[f"normalize_p{pn}"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("model,D,ln", [("complex", 64, 1.0), ("distmult", 32, 1.0), ("simple", 64, 1.0), ("cp", 64, 1.0),
                                        ("rescal", 16, 1.0), ("transe", 64, 1.0), ("transe", 64, 2.0), ("rotate", 64, 1.0)])
def test_x_ns_backward(eng, model, D, ln):
    """Fused negative-sampling backward (S and O slots, positive column included) against the CPU algebra
    (oracle/kge_fold.ns_backward, itself pinned to the reference job's gradients)."""
    from oracle import kge_fold as kf

    E, R, n, K = 501, 5, 37, 150
    ent, rel = orc.make_tables(model, E, R, D, sigma=0.5)
    tri = orc.make_triples(E, R, n)
    g = torch.Generator().manual_seed(3)
    negs = {S: torch.randint(0, E, (n, K), generator=g), O: torch.randint(0, E, (n, K + 7), generator=g)}
    ref_e, ref_r = kf.ns_backward(model, ent.double(), rel.double(), tri, negs, 0.25, ln)
    d_ent, d_rel = eng.x_ns_backward(model, ent.cuda(), rel.cuda(), tri.cuda(), {k: v.cuda() for k, v in negs.items()},
                                     0.25, ln)
    _assert_close(d_ent, ref_e, f"{model} d_ent")
    _assert_close(d_rel, ref_r, f"{model} d_rel")


@pytest.mark.parametrize("model,D", [("complex", 128), ("distmult", 32), ("rescal", 24), ("cp", 64), ("transe", 64), ("rotate", 64)])
@pytest.mark.parametrize("loss", ["bce", "kl"])
def test_x_loss_with_csr_labels(eng, model, D, loss):
    """KvsAll losses with CSR multi-hot labels (duplicates, empty rows, label smoothing) against the oracle on the
    densified label matrix; both directions."""
    E, R, n = 3001, 5, 200
    ent, rel = orc.make_tables(model, E, R, D, sigma=0.3)
    tri = orc.make_triples(E, R, n)
    g = torch.Generator().manual_seed(3)
    counts = (torch.rand((n, E), generator=g) < 0.004).float()
    counts[torch.arange(n), tri[:, O]] += 1.0
    counts[5, int(tri[5, O])] += 1.0          # a duplicate triple: label 2
    counts[7] = 0.0                            # a row without labels
    rows, cols = torch.nonzero(counts, as_tuple=True)
    rep = counts[rows, cols].long()
    cols_rep = torch.repeat_interleave(cols, rep)
    rows_rep = torch.repeat_interleave(rows, rep)
    offs = torch.zeros(n + 1, dtype=torch.int64)
    offs[1:] = torch.cumsum(torch.bincount(rows_rep, minlength=n), 0)
    ce, cr = ent.cuda(), rel.cuda()
    off = 1.0 if loss == "bce" else 0.0
    fn = (lambda x, y: orc.bce_loss(x, y, off)) if loss == "bce" else orc.kl_loss
    dot = model in ("complex", "distmult", "rescal", "cp")
    for combine, qi, sc in (("sp_", tri[:, S], orc.score_sp(model, ent, rel, tri[:, S], tri[:, P])),
                            ("_po", tri[:, O], orc.score_po(model, ent, rel, tri[:, P], tri[:, O]))):
        for eps in ((0.0, 0.1) if dot else (0.0,)):
            lab = orc.kvsall_smooth_labels(counts, eps) if eps > 0 else counts
            ref = float(fn(sc, lab))
            got, rws = eng.x_score_1vsN_loss_csr(model, combine, ce, cr, ce, offs.cuda(), cols_rep.cuda(), qi.cuda(),
                                                 tri[:, P].cuda(), loss, off, eps, return_rows=True)
            assert abs(float(got) - ref) <= 1e-4 * abs(ref), (model, loss, combine, eps, float(got), ref)
            assert abs(float(rws.sum()) - ref) <= 1e-4 * abs(ref)
