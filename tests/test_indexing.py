"""Host-side label plumbing (kge_b200.indexing on the native b200kge_kvsall_* functions) against the live
reference's KvsAllIndex / get_sp_po_coords_from_spo_batch (tests/golden/index.npz).  Integer data: bit-exact.
CPU only."""
import os

import numpy as np
import pytest
import torch

from kge_b200 import indexing

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "index.npz")
S, P, O = 0, 1, 2
KEYS = {"sp": ([S, P], O), "po": ([P, O], S), "so": ([S, O], P)}


@pytest.fixture(scope="module")
def g():
    z = np.load(GOLDEN)
    return {k: torch.from_numpy(z[k]) if z[k].ndim else int(z[k]) for k in z.files}


@pytest.mark.parametrize("key", ["sp", "po", "so"])
def test_index_matches_reference(g, key):
    ix = indexing.index_KvsAll(g["triples"], key)
    assert ix.key_cols == KEYS[key][0] and ix.value_col == KEYS[key][1]
    assert torch.equal(ix._keys, g[f"{key}_keys"]) and ix._keys.dtype == g[f"{key}_keys"].dtype
    assert torch.equal(ix._values_offset, g[f"{key}_offsets"]) and ix._values_offset.dtype == torch.int32
    assert torch.equal(ix._values, g[f"{key}_values"])
    assert len(ix) == g[f"{key}_keys"].shape[0]
    # dictionary-style access
    k0 = tuple(int(x) for x in g[f"{key}_keys"][3])
    lo, hi = int(g[f"{key}_offsets"][3]), int(g[f"{key}_offsets"][4])
    assert torch.equal(ix[k0], g[f"{key}_values"][lo:hi])
    assert ix.get((10 ** 6, 0)) == [] and ix.get((10 ** 6, 0), "x") == "x"
    assert sum(len(v) for v in ix.values()) == g["triples"].shape[0]
    assert list(ix.keys())[3] == k0


def test_get_all_and_label_coords_match_reference(g):
    sp = indexing.index_KvsAll(g["triples"], "sp")
    po = indexing.index_KvsAll(g["triples"], "po")
    batch, E = g["batch"], g["num_entities"]
    assert torch.equal(sp.get_all(batch[:, [S, P]]), g["sp_get_all"])
    assert torch.equal(po.get_all(batch[:, [P, O]]), g["po_get_all"])
    offs, cols = indexing.sp_po_label_csr(batch, E, sp, po)
    assert offs.numel() == batch.shape[0] + 1 and int(offs[-1]) == g["sp_po_coords"].shape[0]
    # same coordinate multiset as the reference (it lists all sp coordinates first, then all po ones)
    mine = indexing.csr_to_coords(offs, cols)
    ref = g["sp_po_coords"].long()
    key = lambda c: (c[:, 0] * (2 * E) + c[:, 1]).sort().values
    assert torch.equal(key(mine), key(ref))
    # densified, both give the label matrix the ranking / KvsAll jobs build (duplicates add up, util.py:46-58)
    dense_ref = torch.sparse_coo_tensor(ref.t(), torch.ones(len(ref)), (batch.shape[0], 2 * E)).to_dense()
    dense_mine = torch.sparse_coo_tensor(mine.t(), torch.ones(len(mine)), (batch.shape[0], 2 * E)).to_dense()
    assert torch.equal(dense_ref, dense_mine)


def test_collate_csr(g):
    ix = indexing.index_KvsAll(g["triples"], "sp")
    ex = torch.tensor([5, 0, len(ix) - 1, 5])
    q, offs, cols = ix.collate_csr(ex)
    assert torch.equal(q, g["sp_keys"][ex].long())
    for i, e in enumerate(ex.tolist()):
        lo, hi = int(g["sp_offsets"][e]), int(g["sp_offsets"][e + 1])
        assert torch.equal(cols[int(offs[i]): int(offs[i + 1])], g["sp_values"][lo:hi].long())
    with pytest.raises(ValueError):
        ix.collate_csr(torch.tensor([len(ix)]))


def test_edge_cases():
    empty = indexing.KvsAllIndex(torch.zeros((0, 3), dtype=torch.int64), [S, P], O)
    assert len(empty) == 0 and empty._values_offset.tolist() == [0]
    offs, cols = empty.get_all_csr(torch.tensor([[1, 2]]))
    assert offs.tolist() == [0, 0] and cols.numel() == 0
    one = indexing.KvsAllIndex(torch.tensor([[7, 1, 3]]), [P, O], S)
    assert one._keys.tolist() == [[1, 3]] and one._values.tolist() == [7]
    with pytest.raises(ValueError):
        indexing.KvsAllIndex(torch.tensor([[7, 1, 3]]), [S, S], O)
    with pytest.raises(ValueError):
        indexing.index_KvsAll(torch.tensor([[7, 1, 3]]), "xx")
    # large random case against a plain Python dictionary
    g = torch.Generator().manual_seed(5)
    tri = torch.randint(0, 50, (5000, 3), generator=g)
    ix = indexing.index_KvsAll(tri, "po")
    d = {}
    for s, p, o in tri.tolist():
        d.setdefault((p, o), []).append(s)
    assert len(ix) == len(d)
    for k, v in list(d.items())[:200]:
        assert ix[k].tolist() == sorted(v)


def test_index_properties_randomised():
    """Property test of the native index against a plain dictionary over random small graphs (many duplicate keys
    and duplicate triples), all three key layouts, lookups of present and absent keys."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.tuples(st.integers(0, 5), st.integers(0, 2), st.integers(0, 5)), min_size=0, max_size=40),
           st.sampled_from(["sp", "po", "so"]),
           st.lists(st.tuples(st.integers(0, 7), st.integers(0, 7)), min_size=0, max_size=10))
    def check(triples, key, queries):
        cols, val = KEYS[key]
        tri = torch.tensor(triples, dtype=torch.int64).view(-1, 3)
        ix = indexing.index_KvsAll(tri, key)
        d = {}
        for t in triples:
            d.setdefault((t[cols[0]], t[cols[1]]), []).append(t[val])
        assert len(ix) == len(d)
        assert [tuple(k) for k in ix._keys.tolist()] == sorted(d)
        assert ix._values.tolist() == [v for k in sorted(d) for v in sorted(d[k])]
        assert ix._values_offset.tolist()[-1] == len(triples)
        q = torch.tensor(queries, dtype=torch.int64).view(-1, 2)
        offs, vals = ix.get_all_csr(q)
        for i, k in enumerate(queries):
            assert vals[int(offs[i]): int(offs[i + 1])].tolist() == sorted(d.get(tuple(k), []))
        if len(ix):
            ex = torch.arange(len(ix) - 1, -1, -1)
            qs, o2, c2 = ix.collate_csr(ex)
            assert [tuple(k) for k in qs.tolist()] == sorted(d)[::-1]
            assert int(o2[-1]) == len(triples)

    check()
