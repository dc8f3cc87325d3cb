"""CPU tests of host-side logic that needs no GPU: argument validation in the Python layer, label
packing, relation-dim rules, model construction errors, the reference arm of bench.py."""
import json
import subprocess
import sys
import os

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_relation_dim_rules():
    from kge_b200.model import relation_dim
    from oracle import kge_oracle as orc

    for m in orc.MODELS:
        for D in (8, 32, 200):
            assert relation_dim(m, D) == orc.relation_dim(m, D)
    assert relation_dim("rescal", 200) == 40000 and relation_dim("rotate", 512) == 256


def test_model_construction_errors_match_reference():
    from kge_b200 import KgeModel, KgeLoss

    for name in ("complex", "simple", "cp", "rotate"):
        with pytest.raises(ValueError, match="even dimensionality"):
            KgeModel(name, 10, 2, 7)                      # simple.py:46-50, cp.py:44-48, rotate.py:87-91
    with pytest.raises(ValueError):
        KgeModel("conve", 10, 2, 8)
    with pytest.raises(ValueError, match="invalid value train.loss"):
        KgeLoss.create("soft_margin")                     # loss.py:87-89
    m = KgeModel("rescal", 10, 3, 6, seed=0)
    assert m.state_dict()["_relation_embedder._embeddings.weight"].shape == (3, 36)
    assert set(m.state_dict()) == {"_entity_embedder._embeddings.weight", "_relation_embedder._embeddings.weight"}
    r = KgeModel("rotate", 10, 3, 8, seed=0)
    w = r._relation_embedder.weight
    assert w.shape == (3, 4) and float(w.abs().max()) <= 3.1416


def test_engine_rejects_bad_inputs_before_touching_the_gpu():
    from kge_b200 import engine

    x = torch.zeros(4, 8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        engine.score_spo("complex", x, x, x)
    with pytest.raises(ValueError):
        engine.score_1vsN("complex", "s_o", x, x, x)


def test_bench_reference_arm_runs_on_cpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "triples/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["higher_is_better"] is True


def test_topk_tie_break_is_lowest_index():
    from kge_b200.sharded import _topk_lowest_index

    v = torch.tensor([[1.0, 3.0, 3.0, 2.0, 3.0]])
    vals, idx = _topk_lowest_index(v, 3)
    assert idx.tolist() == [[1, 2, 4]] and vals.tolist() == [[3.0, 3.0, 3.0]]
    ids = torch.tensor([[9, 4, 7, 1, 2]])
    vals, pos = _topk_lowest_index(v, 2, ids)
    assert torch.gather(ids, 1, pos).tolist() == [[2, 4]]      # among the 3.0s, ids 4,7,2 -> 2 then 4
