"""Metadata filters as a row gate (reference: _build_metadata_filters -> index.as_retriever(filters=...)):
a filtered search must equal the oracle's exact top-k over the ELIGIBLE rows only -- ids (mapped back to corpus rows) and
scores bit-equal -- through every stage-1 variant, the exact gated scan, the repair ladder and the retriever surface."""

import numpy as np
import pytest
import torch

import oracle
from oracle import cport
from tensor_truth_b200 import _lib
from tensor_truth_b200.filters import eligible_rows
from tensor_truth_b200.index import DeviceIndex
from tensor_truth_b200.synth import make_small

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def world():
    tree, bits, inv, q = make_small(30_000, 48, dim=1024, levels=3, seed=5)
    meta = [{"doc_type": ("library", "book", "paper")[r % 3], "year": 2000 + (r * 7) % 25} for r in range(bits.shape[0])]
    return tree, bits, q, meta


def _oracle_filtered(bits, q, k, elig):
    rows = np.nonzero(elig)[0]
    ids, sc, _ = cport.scan_topk(bits[rows], q, k)
    return np.where(ids >= 0, rows[np.clip(ids, 0, len(rows) - 1)], -1), sc


@pytest.mark.parametrize("variant", [_lib.SCAN_TCGEN05, _lib.SCAN_SIMT])
@pytest.mark.parametrize("b", [1, 8, 24, 40])
def test_filtered_search_equals_oracle_on_the_eligible_rows(world, variant, b):
    tree, bits, q, meta = world
    spec = {"doc_type": "paper", "year": {"$gte": 2010}}
    elig = eligible_rows(spec, meta)
    assert 0 < elig.sum() < len(meta) // 3
    idx = DeviceIndex(bits, tree, device=torch.device("cuda:0"), variant=variant)
    rf = idx.row_filter(elig, key="paper>=2010")
    assert idx.row_filter(elig, key="paper>=2010") is rf and rf.n_eligible == int(elig.sum())
    r = idx.search_certified(torch.from_numpy(q[:b]).cuda(), 10, row_filter=rf)
    torch.cuda.synchronize()
    ids_o, sc_o = _oracle_filtered(bits, q[:b], 10, elig)
    got = r.ids.cpu().numpy()
    assert elig[got].all()
    assert (got == ids_o).all() and (r.scores.cpu().numpy() == sc_o).all()
    # unfiltered searches on the same index are untouched
    r0 = idx.search_certified(torch.from_numpy(q[:b]).cuda(), 10)
    torch.cuda.synchronize()
    ids_u, sc_u, _ = cport.scan_topk(bits, q[:b], 10)
    assert (r0.ids.cpu().numpy() == ids_u).all()


def test_exact_gated_scan_and_fewer_eligible_rows_than_k(world):
    tree, bits, q, meta = world
    idx = DeviceIndex(bits, tree, device=torch.device("cuda:0"))
    elig = np.zeros(bits.shape[0], bool)
    elig[[5, 77, 29_999, 12_345]] = True
    rf = idx.row_filter(elig)
    qd = torch.from_numpy(q[:3]).cuda()
    ex = idx.search_exact(qd, 10, row_filter=rf)
    r = idx.search_certified(qd, 10, row_filter=rf)
    torch.cuda.synchronize()
    ids_o, sc_o = _oracle_filtered(bits, q[:3], 10, elig)
    for got in (ex, r):
        g = got.ids.cpu().numpy()
        assert (g[:, :4] == ids_o[:, :4]).all() and (g[:, 4:] == -1).all()
        assert (got.scores.cpu().numpy()[:, :4] == sc_o[:, :4]).all()


def test_retriever_with_filters_auto_merges_the_filtered_list(world):
    from tensor_truth_b200.retriever import B200AutoMergingRetriever, B200VectorIndexRetriever, NodeTable
    from tensor_truth_b200.schema import QueryBundle, TextNode

    tree, bits, q, meta = world
    n_leaf = bits.shape[0]
    nodes = [TextNode(id_=f"n{o}", text="", metadata=dict(meta[o]) if o < n_leaf else {}) for o in range(tree.n_nodes)]
    idx = DeviceIndex(bits, tree, device=torch.device("cuda:0"))
    spec = {"doc_type": ["paper", "book"], "year": {"$lt": 2020}}
    elig = eligible_rows(spec, meta)
    base = B200VectorIndexRetriever(idx, 10, None, NodeTable(nodes=nodes), filters=spec)
    am = B200AutoMergingRetriever(base, None)
    assert base.row_filter is not None and base.row_filter.n_eligible == int(elig.sum())
    merged_any = False
    for b in range(12):
        out = am.retrieve(QueryBundle(query_str="x", embedding=q[b].tolist()))
        ids_o, sc_o = _oracle_filtered(bits, q[b:b + 1], 10, elig)
        exp = oracle.auto_merge([(int(o), float(s)) for o, s in zip(ids_o[0], sc_o[0]) if o >= 0], tree.parent_of, tree.child_count,
                                tree.prev_id, tree.next_id)
        assert [(n.node.id_, n.score) for n in out] == [(f"n{o}", s) for o, s in exp]
        merged_any = merged_any or any(o >= n_leaf for o, _ in exp)
        leaves = base.retrieve(QueryBundle(query_str="x", embedding=q[b].tolist()))
        assert [n.node.id_ for n in leaves] == [f"n{o}" for o in ids_o[0] if o >= 0]
    assert merged_any
