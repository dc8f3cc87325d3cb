"""Pin the CPU oracle against golden vectors produced by the live reference
(tests/golden/gen_golden.py).  CPU only."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import kge_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
S, P, O = 0, 1, 2


def _load(name):
    z = np.load(os.path.join(GOLDEN, name))
    return {k: torch.from_numpy(z[k]) if z[k].ndim else z[k].item() for k in z.files}


def _model_of(tag):
    return tag.split("_")[0]


SCORE_FILES = sorted(os.path.basename(f) for f in glob.glob(os.path.join(GOLDEN, "scores_*.npz")))


def _close(a, b, what):
    # the oracle mirrors the reference's own op sequence: agreement is fp32 round-off
    scale = max(1.0, float(b.abs().max()))
    err = float((a - b).abs().max())
    assert err <= 2e-6 * scale, f"{what}: max|d|={err} scale={scale}"


@pytest.mark.parametrize("fname", SCORE_FILES)
def test_scores_match_reference(fname):
    g = _load(fname)
    model = _model_of(fname[len("scores_"):-4])
    ent, rel, tri, ln = g["ent"], g["rel"], g["triples"], float(g["l_norm"])
    s, p, o = tri[:, S], tri[:, P], tri[:, O]
    _close(orc.score_spo(model, ent, rel, s, p, o, ln), g["spo"], "spo")
    _close(orc.score_sp(model, ent, rel, s, p, None, ln), g["sp"], "sp")
    _close(orc.score_po(model, ent, rel, p, o, None, ln), g["po"], "po")
    _close(orc.score_sp(model, ent, rel, s, p, g["subset"], ln), g["sp_subset"], "sp_subset")
    _close(orc.score_po(model, ent, rel, p, o, g["subset"], ln), g["po_subset"], "po_subset")
    _close(orc.score_so(model, ent, rel, s, o, None, ln), g["so"], "so")
    _close(orc.score_so(model, ent, rel, s, o, g["psub"], ln), g["so_subset"], "so_subset")
    _close(orc.score_sp_po(model, ent, rel, s, p, o, None, ln), g["sp_po"], "sp_po")
    _close(orc.score_sp_po(model, ent, rel, s, p, o, g["subset"], ln), g["sp_po_subset"], "sp_po_subset")


def test_losses_match_reference():
    g = _load("losses.npz")
    x = g["scores"]
    rel = lambda a, b: abs(float(a) - float(b)) <= 2e-6 * max(1.0, abs(float(b)))
    assert rel(orc.bce_loss(x, g["idx"]), g["bce_idx"])
    assert rel(orc.bce_loss(x, g["idx"], 2.0), g["bce_idx_off2"])
    assert rel(orc.bce_loss(x, g["multi"]), g["bce_multi"])
    assert rel(orc.bce_loss(x, g["smooth"]), g["bce_smooth"])
    assert rel(orc.kl_loss(x, g["idx"]), g["kl_idx"])
    assert rel(orc.kl_loss(x, g["multi"]), g["kl_multi"])
    assert rel(orc.kl_loss(x, g["smooth"]), g["kl_smooth"])
    assert rel(orc.kvsall_smooth_labels(g["multi"], 0.1).sum(), g["smooth"].sum())


def test_ranks_match_reference_bit_exact():
    g = _load("ranks.npz")
    r, t = orc.ranks_and_ties(g["sp"], g["o_true"])
    assert torch.equal(r, g["raw_o_rank"]) and torch.equal(t, g["raw_o_ties"])
    r, t = orc.ranks_and_ties(g["po"], g["s_true"])
    assert torch.equal(r, g["raw_s_rank"]) and torch.equal(t, g["raw_s_ties"])
    sr, st, orank, ot = orc.filter_and_rank(g["sp"], g["po"], g["labels"], g["o_true"], g["s_true"])
    assert torch.equal(sr, g["filt_s_rank"]) and torch.equal(st, g["filt_s_ties"])
    assert torch.equal(orank, g["filt_o_rank"]) and torch.equal(ot, g["filt_o_ties"])
    assert torch.equal(orc.final_ranks(orank, ot), g["final_o"])


@pytest.mark.parametrize("model", ["complex", "rotate", "transe", "rescal"])
def test_ns_match_reference(model):
    g = _load(f"ns_{model}.npz")
    ent, rel, tri, ln = g["ent"], g["rel"], g["triples"], float(g["l_norm"])
    _close(orc.score_spo(model, ent, rel, tri[:, S], tri[:, P], tri[:, O], ln), g["pos"], "pos")
    for slot, nm in ((S, "s"), (P, "p"), (O, "o")):
        for impl in ("triple", "batch"):
            got = orc.ns_score(model, ent, rel, tri, g[f"neg_{nm}"], slot, impl, ln)
            _close(got, g[f"ns_{nm}_{impl}"], f"ns {nm} {impl}")
    full = orc.ns_scores_with_positive(model, ent, rel, tri, g["neg_o"], O, "triple", ln)
    _close(full[:, 0], g["pos"], "assembled col0")
    _close(full[:, 1:], g["ns_o_triple"], "assembled negs")


def test_fp64_mode_is_consistent():
    for model in orc.MODELS:
        D = 16 if model == "rescal" else 32
        ent, rel = orc.make_tables(model, 50, 4, D)
        tri = orc.make_triples(50, 4, 8)
        a = orc.score_sp_po(model, ent, rel, tri[:, 0], tri[:, 1], tri[:, 2])
        b = orc.score_sp_po(model, ent.double(), rel.double(), tri[:, 0], tri[:, 1], tri[:, 2])
        rms = float(b.pow(2).mean().sqrt())
        assert float((a.double() - b).abs().max()) <= 1e-5 * max(rms, 1.0)


@pytest.mark.parametrize("model", ["complex", "transe"])
def test_job_traces_match_reference(model):
    """Trace values of the reference's own jobs (TrainingJob1vsAll forward-only epoch and
    EntityRankingJob on the valid split; tests/golden/gen_golden.py:gen_jobs) re-derived by the oracle."""
    g = _load(f"jobs_{model}.npz")
    ent, rel = g["ent"], g["rel"]
    train, valid, test = g["train"].long(), g["valid"].long(), g["test"].long()
    # epoch avg_loss = sum_batches(avg_loss_b * size_b) / num_examples (train.py:405-420,536-548) and
    # avg_loss_b * size_b = loss_sp + loss_po summed over the batch (train_1vsAll.py:48-82): the
    # batch split cancels, leaving one forward over the whole split.
    for loss in ("bce", "kl"):
        got = float(orc.train_1vsall_forward(model, ent, rel, train, loss=loss))
        want = float(g[f"avg_loss_{loss}"])
        assert abs(got - want) <= 1e-5 * abs(want), (loss, got, want)
        batched = sum(float(orc.train_1vsall_forward(model, ent, rel, train[i:i + 16], loss=loss)) * len(train[i:i + 16])
                      for i in range(0, len(train), 16)) / len(train)
        assert abs(batched - want) <= 1e-5 * abs(want), (loss, batched, want)
    # default eval.filter_splits = [train, valid] (+ test only for *_filtered_with_test)
    met = orc.entity_ranking_metrics(model, ent, rel, valid, [train, valid], test_triples=test)
    for k, v in met.items():
        want = float(g["valid_" + k])
        assert abs(v - want) <= 1e-6 * max(1.0, abs(want)), (k, v, want)


@pytest.mark.parametrize("model", ["complex", "transe"])
def test_kvsall_epoch_matches_reference(model):
    """TrainingJobKvsAll forward-only epoch (train_KvsAll.py:205-300): one example per distinct (s,p) and per
    distinct (p,o) of the training split, multi-hot labels (duplicate triples add up, util.py:46-58), optional
    label smoothing, loss summed over examples and divided by their number."""
    g = _load(f"jobs_{model}.npz")
    ent, rel, train = g["ent"], g["rel"], g["train"].long()
    E = ent.shape[0]

    def examples(key_cols, val_col):
        keys, inv = torch.unique(train[:, key_cols], dim=0, return_inverse=True)
        labels = torch.zeros((keys.shape[0], E))
        labels.index_put_((inv, train[:, val_col]), torch.ones(len(train)), accumulate=True)
        return keys, labels

    sp_keys, sp_lab = examples([S, P], O)
    po_keys, po_lab = examples([P, O], S)
    sc_sp = orc.score_sp(model, ent, rel, sp_keys[:, 0], sp_keys[:, 1])
    sc_po = orc.score_po(model, ent, rel, po_keys[:, 0], po_keys[:, 1])
    n = sp_keys.shape[0] + po_keys.shape[0]
    for loss, eps in (("kl", 0.0), ("kl", 0.2), ("bce", 0.2)):
        fn = orc.kl_loss if loss == "kl" else orc.bce_loss
        lab = (lambda y: orc.kvsall_smooth_labels(y, eps)) if eps > 0 else (lambda y: y)
        got = float(fn(sc_sp, lab(sp_lab)) + fn(sc_po, lab(po_lab))) / n
        want = float(g[f"kvsall_avg_loss_{loss}_{int(eps * 10)}"])
        assert abs(got - want) <= 2e-5 * abs(want), (loss, eps, got, want)


@pytest.mark.parametrize("model", ["complex", "rotate"])
def test_negative_sampling_batch_matches_reference_job(model):
    """One batch of TrainingJobNegativeSampling with the sampled negatives replayed: per slot the [n, 1+K]
    score block (positive first), BCE with offset against the first-column labels, divided by the batch size
    (train_negative_sampling.py:113-163)."""
    g = _load(f"nsjob_{model}.npz")
    ent, rel, tri = g["ent"], g["rel"], g["triples"].long()
    n, off = tri.shape[0], float(g["offset"])
    total = 0.0
    for slot, nm in ((S, "s"), (P, "p"), (O, "o")):
        neg = g[f"neg_{nm}"].long()
        for impl in ("triple", "batch"):
            scores = orc.ns_scores_with_positive(model, ent, rel, tri, neg, slot, impl)
            loss = float(orc.bce_loss(scores, orc.ns_labels(n, neg.shape[1]), off)) / n
            if impl == "triple":
                total += loss
            else:
                assert abs(loss - float(orc.bce_loss(orc.ns_scores_with_positive(model, ent, rel, tri, neg, slot, "triple"),
                                                      orc.ns_labels(n, neg.shape[1]), off)) / n) <= 1e-5 * abs(loss)
    assert int(g["size"]) == n
    assert abs(total - float(g["avg_loss"])) <= 2e-5 * abs(float(g["avg_loss"])), (total, float(g["avg_loss"]))


def test_penalties_and_normalisation_match_reference():
    """KgeModel.penalty(batch) for Lp / N3, weighted / unweighted (kge_model.py:603-649, lookup_embedder.py:123-177)
    and the row normalisation hook (lookup_embedder.py:64-69)."""
    import sys
    sys.path.insert(0, GOLDEN)
    from gen_golden import PENALTY_CASES      # the option sets the fixtures were generated with

    g = _load("penalties.npz")
    for tag, model, eo, ro in PENALTY_CASES:
        space = "complex" if model == "complex" else "euclidean"
        got = float(orc.model_penalty(g[f"{tag}_ent"], g[f"{tag}_rel"], g[f"{tag}_triples"].long(),
                                      dict(eo, space=space), dict(ro, space=space)))
        want = float(g[f"{tag}_total"])
        assert abs(got - want) <= 1e-5 * abs(want), (tag, got, want)
    for pn in (1, 2):
        got = orc.normalize_embeddings(g["normalize_in"], float(pn))
        assert torch.allclose(got, g[f"normalize_p{pn}"], rtol=1e-6, atol=1e-7)
        assert torch.allclose(got.abs().pow(pn).sum(1).pow(1.0 / pn), torch.ones(got.shape[0]), atol=1e-5)


@pytest.mark.parametrize("base", ["complex", "transe"])
def test_reciprocal_relations_model_matches_reference(base):
    g = _load(f"reciprocal_{base}.npz")
    ent, rel2, tri, R, sub = g["ent"], g["rel2"], g["triples"].long(), int(g["num_relations"]), g["subset"].long()
    s, p, o = tri[:, S], tri[:, P], tri[:, O]
    _close(orc.reciprocal_score_spo(base, ent, rel2, s, p, o, "o", R), g["spo_o"], "spo o")
    _close(orc.reciprocal_score_spo(base, ent, rel2, s, p, o, "s", R), g["spo_s"], "spo s")
    _close(orc.score_sp(base, ent, rel2, s, p), g["sp"], "sp")
    _close(orc.reciprocal_score_po(base, ent, rel2, p, o, R), g["po"], "po")
    _close(orc.reciprocal_score_po(base, ent, rel2, p, o, R, sub), g["po_subset"], "po subset")
    _close(orc.reciprocal_score_sp_po(base, ent, rel2, s, p, o, R), g["sp_po"], "sp_po")
    _close(orc.reciprocal_score_sp_po(base, ent, rel2, s, p, o, R, sub), g["sp_po_subset"], "sp_po subset")
    with pytest.raises(Exception, match="undirected"):
        orc.reciprocal_score_spo(base, ent, rel2, s, p, o, None, R)


@pytest.mark.parametrize("model", orc.MODELS)
def test_mid_size_scores_match_reference(model):
    """Mid-size live-reference goldens (E=5003, D=128, n=300: several K chunks and tiles of the tensor-core
    kernels): sampled score columns, row sums and row-wise scores pin the oracle at a shape where the CUDA path
    runs its full pipeline (tests/test_gpu_parity.py::test_mid_size_golden replays them on the GPU)."""
    z = np.load(os.path.join(GOLDEN, f"mid_{model}.npz"))
    E, R, D, n = int(z["E"]), int(z["R"]), int(z["D"]), int(z["n"])
    ent, rel = orc.make_tables(model, E, R, D, sigma=0.5, seed=4321)
    tri = orc.make_triples(E, R, n, seed=17)
    cols = torch.from_numpy(z["cols"])
    sp = orc.score_sp(model, ent, rel, tri[:, 0], tri[:, 1])
    po = orc.score_po(model, ent, rel, tri[:, 1], tri[:, 2])
    for got, key in ((sp, "sp"), (po, "po")):
        rms = float(z[key + "_rms"])
        assert float((got[:, cols] - torch.from_numpy(z[key + "_cols"])).abs().max()) <= 2e-6 * rms
        assert float((got.double().sum(1) - torch.from_numpy(z[key + "_rowsum"])).abs().max()) <= 2e-6 * rms * E ** 0.5
    spo = orc.score_spo(model, ent, rel, tri[:, 0], tri[:, 1], tri[:, 2])
    assert float((spo - torch.from_numpy(z["spo"])).abs().max()) <= 2e-6 * float(z["sp_rms"])
