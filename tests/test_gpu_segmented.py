"""Multi-index as one segmented corpus: one pass of the device pipeline must give, per index, exactly
what that index alone gives -- and the combined, balanced list must be what the reference's MultiIndexRetriever
(restated in oracle/multi_index.py, pinned against the reference in tests/test_reference_pinning.py) makes of the
per-index B200 retrievers.

Every test here needs a GPU:  python -m pytest tests -m gpu
"""

import numpy as np
import pytest
import torch

import oracle
from oracle import cport
from oracle.multi_index import MultiIndexRetriever
from tensor_truth_b200.index import DeviceIndex
from tensor_truth_b200.retriever import B200AutoMergingRetriever, B200MultiIndexRetriever, B200VectorIndexRetriever
from tensor_truth_b200.schema import QueryBundle
from tensor_truth_b200.segmented import SegmentedIndex
from tensor_truth_b200.synth import make_small

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _np(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def parts():
    """Five indexes of very different sizes (one smaller than k, boundaries off the 128-row tile grid)."""
    sizes = [(30_000, 3, 11), (777, 3, 12), (5, 2, 13), (12_345, 4, 14), (64_000, 3, 15)]
    out = []
    for n, levels, seed in sizes:
        tree, bits, inv, q = make_small(n, 8, dim=1024, levels=levels, seed=seed)
        out.append((tree, bits, q))
    return out


def _queries(parts):
    # queries aimed at different segments, so that every index has near hits for some of them
    return np.concatenate([p[2][:2] for p in parts], axis=0).astype(np.float32)


def test_per_segment_topk_equals_each_index_alone(parts):
    seg = SegmentedIndex([p[1] for p in parts], [p[0] for p in parts], device=DEV)
    q = _queries(parts)
    qd = torch.from_numpy(q).to(DEV)
    k = 10
    for b in (1, 3, 10):  # 1 pass, 1 pass, 2 passes of 8 queries
        r = seg.search_certified(qd[:b], k)
        torch.cuda.synchronize()
        ids, scores = _np(r.ids).reshape(seg.n_seg, b, k), _np(r.scores).reshape(seg.n_seg, b, k)
        for s, (tree, bits, _) in enumerate(parts):
            ids_o, sc_o, _ = cport.scan_topk(bits, q[:b], k)
            want = np.where(ids_o >= 0, ids_o + seg.leaf_off[s], -1)
            assert (ids[s] == want).all(), (b, s)
            assert (scores[s] == sc_o).all(), (b, s)
    assert seg.fallbacks == 0


def test_segmented_retrieve_host_equals_per_index_retrieve(parts):
    seg = SegmentedIndex([p[1] for p in parts], [p[0] for p in parts], device=DEV)
    q = _queries(parts)
    for rep in range(6):  # crosses the CUDA-graph capture threshold
        ids, scores, lens = seg.retrieve_host(torch.from_numpy(q[rep:rep + 1]), 10)
        for s, (tree, bits, _) in enumerate(parts):
            exp = oracle.retrieve(bits, q[rep], 10, tree)
            got = [(seg.segment_of(int(o)), float(sc)) for o, sc in zip(ids[s, 0, :lens[s, 0]], scores[s, 0, :lens[s, 0]])]
            assert [(g[0][1], g[1]) for g in got] == exp, (rep, s)
            assert all(g[0][0] == s for g in got)
    assert seg._ws[("graph", 1, 10, 0.5, True, 0)]["graph"] is not None


@pytest.mark.parametrize("strategy", ["top_k_per_index", "none"])
def test_multi_index_retriever_equals_reference_caller_over_per_index_retrievers(parts, strategy):
    """B200MultiIndexRetriever(one segmented index) == MultiIndexRetriever([one B200 retriever per index])."""
    q = _queries(parts)
    seg = SegmentedIndex([p[1] for p in parts], [p[0] for p in parts], device=DEV)
    multi = B200MultiIndexRetriever(seg, similarity_top_k=10, balance_strategy=strategy)
    singles = [B200AutoMergingRetriever(B200VectorIndexRetriever(DeviceIndex(bits, tree, device=DEV), 10), None)
               for tree, bits, _ in parts]

    class Fixed:  # the caller passes a bundle without an embedding (rag_engine.py:416): every child embeds the string
        def __init__(self, vec):
            self.vec = vec

        def get_agg_embedding_from_queries(self, strs):
            return self.vec

    for i in range(6):
        emb = Fixed(q[i].tolist())
        for r in singles:
            r._vector_retriever.embed_model = emb
        multi._embedder.embed_model = emb
        ref = MultiIndexRetriever(singles, balance_strategy=strategy)
        want = [(n.node.metadata["_source_index"], n.node.id_, n.score) for n in ref.retrieve(f"query {i}")]
        got = [(n.node.metadata["_source_index"], n.node.id_, n.score) for n in multi.retrieve(f"query {i}")]
        if strategy == "none":  # the reference's order is thread-completion order
            want, got = sorted(want), sorted(got)
        assert got == want, i
    # LRU by query string, like the reference
    first = multi.retrieve("query 5")
    assert multi.retrieve("query 5") is first
    multi.clear_cache()
    assert multi.retrieve("query 5") is not first
    # a bundle that carries its embedding is used as is
    out = multi.retrieve(QueryBundle(query_str="with embedding", embedding=q[0].tolist()))
    emb0 = Fixed(q[0].tolist())
    multi._embedder.embed_model = emb0
    assert [(n.node.id_, n.score) for n in out] == [(n.node.id_, n.score) for n in multi.retrieve("same vector, by string")]


def test_near_duplicate_segment_is_repaired_by_that_segments_exact_scan():
    rng = np.random.default_rng(5)
    base = rng.standard_normal(1024).astype(np.float32)
    dup = oracle.f32_to_bf16_bits(base[None, :] * (1.0 + 2e-3 * rng.standard_normal((9_000, 1024)).astype(np.float32)))
    tree, bits, inv, q = make_small(20_000, 2, dim=1024, levels=3, seed=3)
    seg = SegmentedIndex([bits, dup], None, device=DEV)
    qq = np.stack([q[0], (base * (1.0 + 1e-3 * rng.standard_normal(1024))).astype(np.float32)])
    r = seg.search_certified(torch.from_numpy(qq).to(DEV), 10)
    torch.cuda.synchronize()
    ids, scores = _np(r.ids).reshape(2, 2, 10), _np(r.scores).reshape(2, 2, 10)
    for s, c in enumerate((bits, dup)):
        ids_o, sc_o, _ = cport.scan_topk(c, qq, 10)
        assert (ids[s] == ids_o + seg.leaf_off[s]).all() and (scores[s] == sc_o).all()
    assert seg.fallbacks >= 1
