"""The algebra the CUDA path is built on (oracle/kge_fold.py): folded 1-vs-N scoring for all seven models, and
the analytic backward of the fused score+loss step (dot family) that the gradient kernels will implement —
against the oracle's scores, autograd, and table gradients recorded from the live reference
(tests/golden/grads_*.npz).  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import kge_fold as kf
from oracle import kge_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
S, P, O = 0, 1, 2
DOT = ("complex", "distmult", "simple", "cp", "rescal")


def _load(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(z[k]) if z[k].ndim else z[k].item() for k in z.files}


@pytest.mark.parametrize("model", orc.MODELS)
@pytest.mark.parametrize("l_norm", [1.0, 2.0])
def test_folded_scores_equal_oracle_scores(model, l_norm):
    if model not in ("transe", "rotate") and l_norm != 1.0:
        pytest.skip("l_norm only matters for the distance family")
    E, R, D, n = 61, 5, 8 if model == "rescal" else 16, 9
    ent, rel = orc.make_tables(model, E, R, D, dtype=torch.float64)
    tri = orc.make_triples(E, R, n)
    s, p, o = tri[:, S], tri[:, P], tri[:, O]
    assert torch.allclose(kf.score_1vsN(model, "sp_", ent, rel, s, p, l_norm), orc.score_sp(model, ent, rel, s, p, l_norm=l_norm), rtol=1e-10, atol=1e-10)
    assert torch.allclose(kf.score_1vsN(model, "_po", ent, rel, o, p, l_norm), orc.score_po(model, ent, rel, p, o, l_norm=l_norm), rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("model", orc.MODELS)
@pytest.mark.parametrize("combine", ["sp_", "_po"])
def test_unfold_is_the_vjp_of_fold(model, combine):
    g = torch.Generator().manual_seed(3)
    n, D = 7, 6 if model == "rescal" else 12
    a = torch.randn((n, D), generator=g, dtype=torch.float64, requires_grad=True)
    p = torch.randn((n, orc.relation_dim(model, D)), generator=g, dtype=torch.float64, requires_grad=True)
    Q = kf.fold(model, combine, a, p)
    dQ = torch.randn(Q.shape, generator=g, dtype=torch.float64)
    ga, gp = torch.autograd.grad(Q, (a, p), dQ)
    da, dp = kf.unfold(model, combine, a.detach(), p.detach(), dQ)
    assert torch.allclose(da, ga, rtol=1e-12, atol=1e-12) and torch.allclose(dp, gp, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("model", orc.MODELS)
@pytest.mark.parametrize("loss", ["bce", "kl"])
def test_analytic_backward_equals_autograd_of_the_oracle_step(model, loss):
    E, R, D, n = 47, 4, 6 if model == "rescal" else 12, 11
    ent, rel = orc.make_tables(model, E, R, D, sigma=0.7, dtype=torch.float64)
    tri = orc.make_triples(E, R, n)
    tri[3] = tri[2]                                    # duplicate rows: scatter-add must accumulate
    off = 0.75 if loss == "bce" else 0.0
    e, r = ent.clone().requires_grad_(True), rel.clone().requires_grad_(True)
    orc.train_1vsall_forward(model, e, r, tri, loss, off).backward()
    d_ent, d_rel = kf.train_1vsall_backward(model, ent, rel, tri, loss, off)
    assert torch.allclose(d_ent, e.grad, rtol=1e-9, atol=1e-12)
    assert torch.allclose(d_rel, r.grad, rtol=1e-9, atol=1e-12)


GRAD_FILES = sorted(os.path.basename(f) for f in glob.glob(os.path.join(GOLDEN, "grads_*.npz")))


@pytest.mark.parametrize("fname", GRAD_FILES)
def test_gradients_match_the_live_reference(fname):
    """Autograd of the oracle's step (all models) and the analytic assembly (dot family) reproduce the table
    gradients of the reference's own backward."""
    g = _load(fname)
    model, loss = fname[len("grads_"):-4].split("_")
    ent, rel, tri, off = g["ent"], g["rel"], g["triples"], float(g["offset"])
    e, r = ent.clone().requires_grad_(True), rel.clone().requires_grad_(True)
    l = orc.train_1vsall_forward(model, e, r, tri, loss, off)
    l.backward()
    assert l.item() == pytest.approx(float(g["loss"]), rel=1e-5)

    def close(a, b, what):
        scale = max(float(b.abs().max()), 1e-6)
        assert float((a - b).abs().max()) <= 2e-5 * scale, (fname, what, float((a - b).abs().max()), scale)

    close(e.grad, g["d_ent"], "autograd d_ent")
    close(r.grad, g["d_rel"], "autograd d_rel")
    d_ent, d_rel = kf.train_1vsall_backward(model, ent, rel, tri, loss, off)
    close(d_ent, g["d_ent"], "analytic d_ent")
    close(d_rel, g["d_rel"], "analytic d_rel")


@pytest.mark.parametrize("loss", ["bce", "kl"])
def test_loss_grad_with_label_matrices(loss):
    """dL/dz for multi-hot and smoothed label matrices (KvsAll) against autograd of the oracle's losses; KL
    normalises the label rows first (loss.py:209-213), rows without labels contribute nothing."""
    g = torch.Generator().manual_seed(5)
    n, E = 6, 29
    z = (torch.randn((n, E), generator=g, dtype=torch.float64) * 2).requires_grad_(True)
    multi = (torch.rand((n, E), generator=g) < 0.15).double()
    multi[0] = 0.0
    multi[1, 3] = 2.0                                   # duplicate triple: label 2
    for lab in (multi, orc.kvsall_smooth_labels(multi, 0.1)):
        off = 0.3 if loss == "bce" else 0.0
        l = (orc.bce_loss(z, lab, off) if loss == "bce" else orc.kl_loss(z, lab)) / n
        (ga,) = torch.autograd.grad(l, z)
        assert torch.allclose(kf.loss_grad(z.detach(), lab, loss, off, n), ga, rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("model", ["complex", "rotate"])
def test_negative_sampling_backward_matches_reference_job(model):
    """Row-wise backward (positives + negatives of all three slots) against the table gradients of the reference's
    own TrainingJobNegativeSampling batch (tests/golden/nsjob_*.npz) and against autograd of the oracle."""
    g = _load(f"nsjob_{model}.npz")
    ent, rel, tri, off = g["ent"], g["rel"], g["triples"].long(), float(g["offset"])
    negs = {S: g["neg_s"], P: g["neg_p"], O: g["neg_o"]}
    d_ent, d_rel = kf.ns_backward(model, ent, rel, tri, negs, off)

    def close(a, b, what):
        scale = max(float(b.abs().max()), 1e-6)
        assert float((a - b).abs().max()) <= 5e-5 * scale, (model, what, float((a - b).abs().max()), scale)

    close(d_ent, g["d_ent"], "d_ent vs reference")
    close(d_rel, g["d_rel"], "d_rel vs reference")
    e, r = ent.clone().requires_grad_(True), rel.clone().requires_grad_(True)
    n = tri.shape[0]
    total = 0.0
    for slot, neg in negs.items():
        scores = orc.ns_scores_with_positive(model, e, r, tri, neg.long(), slot, "triple")
        total = total + orc.bce_loss(scores, orc.ns_labels(n, neg.shape[1]), off) / n
    total.backward()
    close(d_ent, e.grad, "d_ent vs autograd")
    close(d_rel, r.grad, "d_rel vs autograd")
