"""Host-side logic on CPU: the caller restatement (MultiIndexRetriever plumbing, mirroring the reference's own
tests/unit/test_rag_engine.py:17-245), the docstore -> flat-array importer core, carrier types."""

from types import SimpleNamespace
from unittest.mock import MagicMock

import os

import numpy as np
import pytest

from oracle.multi_index import MultiIndexRetriever
from tensor_truth_b200.schema import NodeWithScore, QueryBundle, TextNode
from tensor_truth_b200.tree import build_uniform_tree, tree_from_relations


def _mock_retriever(nodes):
    r = MagicMock()
    r.retrieve.return_value = nodes
    return r


def test_multi_index_combines_and_skips_empty_and_failing():
    n1, n2 = MagicMock(score=0.9), MagicMock(score=0.8)
    m = MultiIndexRetriever([_mock_retriever([n1]), _mock_retriever([n2])])
    out = m._retrieve(QueryBundle(query_str="q"))
    assert len(out) == 2 and n1 in out and n2 in out
    m = MultiIndexRetriever([_mock_retriever([n1]), _mock_retriever([])])
    assert m._retrieve(QueryBundle(query_str="q")) == [n1]
    bad = MagicMock()
    bad.retrieve.side_effect = RuntimeError("libtt_b200 error -2: boom")
    m = MultiIndexRetriever([_mock_retriever([n1]), bad])
    assert m._retrieve(QueryBundle(query_str="q")) == [n1]


def test_multi_index_order_without_balancing_and_cache():
    nodes = [MagicMock() for _ in range(5)]
    m = MultiIndexRetriever([_mock_retriever(nodes[:3]), _mock_retriever(nodes[3:])], balance_strategy="none")
    out = m._retrieve(QueryBundle(query_str="q"))
    assert sorted(map(id, out)) == sorted(map(id, nodes)) and len(out) == 5
    r = _mock_retriever([MagicMock()])
    m = MultiIndexRetriever([r], enable_cache=True)
    m._retrieve(QueryBundle(query_str="q"))
    m._retrieve(QueryBundle(query_str="q"))
    assert r.retrieve.call_count == 1 and m._retrieve_cached.cache_info().currsize == 1
    m.clear_cache()
    assert m._retrieve_cached.cache_info().currsize == 0
    m2 = MultiIndexRetriever([r], enable_cache=False)
    m2.clear_cache()
    m2.clear_cache()


def test_multi_index_balancing_tags_and_truncates():
    a = [NodeWithScore(TextNode(id_=f"a{i}"), s) for i, s in enumerate((0.9, 0.8, 0.7, 0.65))]
    b = [NodeWithScore(TextNode(id_=f"b{i}"), s) for i, s in enumerate((0.6, 0.5))]
    m = MultiIndexRetriever([_mock_retriever(a), _mock_retriever(b)])
    out = m._retrieve(QueryBundle(query_str="q"))
    # 6 nodes / 2 indexes -> 3 per index at most; index 1 only has 2
    assert [n.node.id_ for n in out] == ["a0", "a1", "a2", "b0", "b1"]
    assert [n.node.metadata["_source_index"] for n in out] == [0, 0, 0, 1, 1]
    assert [n.score for n in out] == sorted((n.score for n in out), reverse=True)


def test_multi_index_passes_a_bundle_without_embedding():
    seen = []

    class R:
        def retrieve(self, qb):
            seen.append((qb.query_str, qb.embedding))
            return []

    MultiIndexRetriever([R(), R()]).retrieve("what is a tensor")
    assert seen == [("what is a tensor", None)] * 2


# --------------------------------------------------------------------------- importer core
def _as_docstore(tree):
    """What docstore.json holds for a tree (ids are strings, relations by id), leaves in corpus-row order."""
    nid = lambda o: f"n{o:06d}"  # noqa: E731
    parent, prev, nxt, children = {}, {}, {}, {}
    for o in range(tree.n_nodes):
        parent[nid(o)] = nid(tree.parent_of[o]) if tree.parent_of[o] >= 0 else None
        prev[nid(o)] = nid(tree.prev_id[o]) if tree.prev_id[o] >= 0 else None
        nxt[nid(o)] = nid(tree.next_id[o]) if tree.next_id[o] >= 0 else None
        children[nid(o)] = []
    for o in range(tree.n_nodes):
        if tree.parent_of[o] >= 0:
            children[nid(tree.parent_of[o])].append(nid(o))
    return [nid(o) for o in range(tree.n_nodes)], parent, children, prev, nxt, [nid(o) for o in range(tree.n_leaf)]


@pytest.mark.parametrize("levels", [1, 3, 4])
def test_tree_from_relations_round_trip(levels):
    t = build_uniform_tree(500, levels=levels, seed=5)
    ids, parent, children, prev, nxt, leaf_order = _as_docstore(t)
    rng = np.random.default_rng(0)
    shuffled = [ids[i] for i in rng.permutation(len(ids))]  # docstore order is arbitrary
    t2 = tree_from_relations(shuffled, parent, children, prev, nxt, leaf_order)
    t2.validate()
    assert t2.n_leaf == 500 and t2.node_ids[:500] == leaf_order
    pos = {nid: o for o, nid in enumerate(t2.node_ids)}
    for o in range(t.n_nodes):
        o2 = pos[f"n{o:06d}"]
        assert t2.child_count[o2] == t.child_count[o]
        for a, b in ((t.parent_of, t2.parent_of), (t.prev_id, t2.prev_id), (t.next_id, t2.next_id)):
            assert (b[o2] == -1) == (a[o] == -1)
            if a[o] >= 0:
                assert t2.node_ids[b[o2]] == f"n{a[o]:06d}"
    # leaves come first, then deeper internal levels before shallower ones
    depth = lambda o: 0 if t2.parent_of[o] < 0 else 1 + depth(t2.parent_of[o])  # noqa: E731
    d = [depth(o) for o in range(t2.n_leaf, t2.n_nodes)]
    assert d == sorted(d, reverse=True)


def test_carrier_types_expose_what_consumers_read():
    n = TextNode(id_="abc", text="hello", metadata={"file_name": "x.md"}, parent_id="p", prev_id=None, next_id="nx", child_ids=[])
    ns = NodeWithScore(node=n, score=0.5)
    assert ns.node.id_ == ns.node.node_id == ns.node_id == "abc"
    assert ns.get_score() == 0.5 and NodeWithScore(node=n).get_score() == 0.0
    assert ns.get_content() == ns.text == "hello"
    ns.node.metadata["_source_index"] = 3
    assert ns.metadata["_source_index"] == 3
    assert n.parent_node.node_id == "p" and n.prev_node is None and n.next_node.node_id == "nx" and n.child_nodes is None
    with pytest.raises(ValueError):
        NodeWithScore(node=n).get_score(raise_error=True)
    qb = QueryBundle(query_str="q")
    assert qb.embedding is None and qb.embedding_strs == ["q"]


# --------------------------------------------------------------------------- importer (duck-typed docstore / collection)
def _fake_docstore(tree):
    from tensor_truth_b200.schema import TextNode

    nid = lambda o: f"uuid-{o:05d}"  # noqa: E731
    docs = {}
    kids = {o: [] for o in range(tree.n_nodes)}
    for o in range(tree.n_nodes):
        if tree.parent_of[o] >= 0:
            kids[int(tree.parent_of[o])].append(o)
    for o in range(tree.n_nodes):
        docs[nid(o)] = TextNode(id_=nid(o), text=f"text {o}", metadata={},
                                parent_id=nid(tree.parent_of[o]) if tree.parent_of[o] >= 0 else None,
                                prev_id=nid(tree.prev_id[o]) if tree.prev_id[o] >= 0 else None,
                                next_id=nid(tree.next_id[o]) if tree.next_id[o] >= 0 else None,
                                child_ids=[nid(c) for c in kids[o]])
    return docs


def test_importer_flattens_docstore_and_collection():
    from tensor_truth_b200.importer import flatten_index

    t = build_uniform_tree(300, levels=3, seed=2)
    docs = _fake_docstore(t)
    rng = np.random.default_rng(1)
    order = rng.permutation(300)  # the vector store returns leaves in its own order
    leaf_ids = [f"uuid-{o:05d}" for o in order]
    emb = rng.standard_normal((300, 16)).astype(np.float32)
    corpus, tree, nodes = flatten_index(leaf_ids, emb.tolist(), docs)
    assert corpus.dtype == np.float32 and (corpus == emb).all()
    assert [n.id_ for n in nodes[:300]] == leaf_ids and len(nodes) == t.n_nodes
    # relations survive the re-ordering: check through ids
    for o2 in range(tree.n_nodes):
        n = nodes[o2]
        assert (tree.parent_of[o2] < 0) == (n.parent_id is None)
        if n.parent_id is not None:
            assert nodes[tree.parent_of[o2]].id_ == n.parent_id
        if n.next_id is not None:
            assert nodes[tree.next_id[o2]].id_ == n.next_id
        assert tree.child_count[o2] == len(n.child_ids)
    with pytest.raises(ValueError):
        flatten_index(leaf_ids + ["ghost"], emb.tolist() + [[0.0] * 16], docs)
    # a docstore leaf WITHOUT an embedding does not fail the load (the reference would still serve the index): it is
    # warned about, gets an ordinal after the embedded leaves and still counts among its parent's children
    with pytest.warns(UserWarning, match="no embedding"):
        corpus2, tree2, nodes2 = flatten_index(leaf_ids[:-1], emb[:-1].tolist(), docs)
    assert corpus2.shape[0] == 299 and tree2.n_leaf == 299 and tree2.n_nodes == t.n_nodes
    stray = [o for o in range(tree2.n_leaf, tree2.n_nodes) if nodes2[o].id_ == leaf_ids[-1]]
    assert len(stray) == 1 and tree2.child_count[stray[0]] == 0
    assert tree2.child_count[tree2.parent_of[stray[0]]] == len(nodes2[tree2.parent_of[stray[0]]].child_ids)


def _persisted_docstore(tree):
    """``docstore.json`` as upstream's SimpleDocumentStore persists it [U]: nodes wrapped in {"__data__", "__type__"},
    relationships keyed by the NodeRelationship VALUE ("1" source ... "5" child), CHILD a list, the others single."""
    nid = lambda o: f"uuid-{o:05d}"  # noqa: E731

    def rel(o, node_type="1"):
        return {"node_id": nid(o), "node_type": node_type, "metadata": {}, "hash": f"h{o}", "class_name": "RelatedNodeInfo"}

    kids = {o: [] for o in range(tree.n_nodes)}
    for o in range(tree.n_nodes):
        if tree.parent_of[o] >= 0:
            kids[int(tree.parent_of[o])].append(o)
    data = {}
    for o in range(tree.n_nodes):
        r = {"1": {"node_id": "doc-0", "node_type": "4", "metadata": {}, "hash": "d", "class_name": "RelatedNodeInfo"}}
        if tree.prev_id[o] >= 0:
            r["2"] = rel(int(tree.prev_id[o]))
        if tree.next_id[o] >= 0:
            r["3"] = rel(int(tree.next_id[o]))
        if tree.parent_of[o] >= 0:
            r["4"] = rel(int(tree.parent_of[o]))
        if kids[o]:
            r["5"] = [rel(c) for c in kids[o]]
        data[nid(o)] = {"__type__": "1",
                        "__data__": {"id_": nid(o), "embedding": None, "metadata": {"file_name": "a.md", "page": o % 7},
                                     "excluded_embed_metadata_keys": [], "excluded_llm_metadata_keys": [], "relationships": r,
                                     "text": f"text {o}", "mimetype": "text/plain", "start_char_idx": 0, "end_char_idx": 6,
                                     "text_template": "{metadata_str}\n\n{content}", "metadata_template": "{key}: {value}",
                                     "metadata_seperator": "\n", "class_name": "TextNode"}}
    return {"docstore/metadata": {k: {"doc_hash": "x", "ref_doc_id": "doc-0"} for k in data},
            "docstore/data": data, "docstore/ref_doc_info": {"doc-0": {"node_ids": list(data), "metadata": {}}}}


def test_importer_reads_the_persisted_docstore_json(tmp_path):
    """The relations come out of ``docstore.json`` itself -- no llama_index -- and give the same flat tree as the loaded
    docstore's node objects do (``flatten_index`` over either)."""
    import json

    from tensor_truth_b200.importer import StoredNode, flatten_index, load_docstore_json

    t = build_uniform_tree(300, levels=3, seed=2)
    index_dir = tmp_path / "index"
    index_dir.mkdir()
    (index_dir / "docstore.json").write_text(json.dumps(_persisted_docstore(t)))
    stored = load_docstore_json(str(index_dir))            # the directory, as document_index.py:138 names the file
    assert stored.keys() == load_docstore_json(str(index_dir / "docstore.json")).keys()
    assert len(stored) == t.n_nodes and all(isinstance(n, StoredNode) for n in stored.values())
    n0 = stored["uuid-00000"]
    assert n0.node_id == "uuid-00000" and n0.get_content() == "text 0" and n0.metadata == {"file_name": "a.md", "page": 0}
    assert n0.source_node.node_id == "doc-0" and n0.parent_node is not None and n0.child_nodes is None
    rng = np.random.default_rng(1)
    order = rng.permutation(300)
    leaf_ids = [f"uuid-{o:05d}" for o in order]
    emb = rng.standard_normal((300, 16)).astype(np.float32)
    corpus_a, tree_a, nodes_a = flatten_index(leaf_ids, emb, stored)
    corpus_b, tree_b, nodes_b = flatten_index(leaf_ids, emb, _fake_docstore(t))
    for name in ("parent_of", "child_count", "prev_id", "next_id"):
        assert (getattr(tree_a, name) == getattr(tree_b, name)).all(), name
    assert tree_a.n_leaf == tree_b.n_leaf == 300 and [n.id_ for n in nodes_a] == [n.id_ for n in nodes_b]
    # a file that is not a SimpleDocumentStore dump is refused by name, not by a KeyError somewhere downstream
    (index_dir / "other.json").write_text("{}")
    with pytest.raises(ValueError, match="docstore/data"):
        load_docstore_json(str(index_dir / "other.json"))


def test_importer_defaults_to_the_score_the_reference_reports():
    """the reference opens Chroma with the default (squared-L2) space, ChromaVectorStore reports
    exp(-distance); the importer must default to that, refuse other spaces, and let the caller override."""
    from tensor_truth_b200 import _lib
    from tensor_truth_b200.importer import collection_score_mode

    class Default:
        metadata = None

    class Cosine:
        metadata = {"hnsw:space": "cosine"}

    assert collection_score_mode(Default()) == _lib.SCORE_CHROMA_L2_EXP
    assert collection_score_mode(object()) == _lib.SCORE_CHROMA_L2_EXP
    with pytest.raises(ValueError):
        collection_score_mode(Cosine())


def test_query_is_embedded_once_across_per_index_retrievers():
    """one QueryBundle fanned out over several retrievers on a thread pool is embedded once."""
    import time
    from concurrent.futures import ThreadPoolExecutor

    from tensor_truth_b200.retriever import B200VectorIndexRetriever, NodeTable

    class Embedder:
        calls = 0

        def get_agg_embedding_from_queries(self, strs):
            Embedder.calls += 1
            time.sleep(0.05)  # a real embed model takes 10-30 ms
            return [0.25] * 16

    dummy_index = SimpleNamespace(tree=None)
    emb = Embedder()
    retrievers = [B200VectorIndexRetriever(dummy_index, 10, emb, NodeTable()) for _ in range(6)]
    qb = QueryBundle(query_str="shared question")
    with ThreadPoolExecutor(max_workers=6) as pool:
        outs = list(pool.map(lambda r: r._query_tensor(qb), retrievers))
    assert Embedder.calls == 1
    assert all(o.shape == (1, 16) and float(o[0, 0]) == 0.25 for o in outs)
    assert qb.embedding == [0.25] * 16


# --------------------------------------------------------------------------- multi-index as one segmented corpus (N4), host side
def test_concat_trees_keeps_every_relation_and_maps_back():
    from tensor_truth_b200.tree import concat_trees, locate

    trees = [build_uniform_tree(50, 3, 1), build_uniform_tree(7, 3, 2), build_uniform_tree(120, 2, 3)]
    t, leaf_off, int_off = concat_trees(trees)
    t.validate()
    assert t.n_leaf == 177 and t.n_nodes == sum(x.n_nodes for x in trees)
    seen = set()
    for o in range(t.n_nodes):
        s, l = locate(o, trees, leaf_off, int_off)
        seen.add((s, l))
        assert (o < t.n_leaf) == (l < trees[s].n_leaf)
        assert t.child_count[o] == trees[s].child_count[l]
        for comb, own in ((t.parent_of, trees[s].parent_of), (t.prev_id, trees[s].prev_id), (t.next_id, trees[s].next_id)):
            if own[l] < 0:
                assert comb[o] == -1
            else:
                assert locate(int(comb[o]), trees, leaf_off, int_off) == (s, int(own[l]))
    assert len(seen) == t.n_nodes


def test_b200_multi_index_retriever_matches_reference_golden_cases():
    """The combining logic of B200MultiIndexRetriever on the scenarios recorded from the reference's own
    MultiIndexRetriever (tests/golden/multi_index_ref.json), with a stand-in for the device index."""
    import json
    import os

    from tensor_truth_b200.retriever import B200MultiIndexRetriever, NodeTable

    with open(os.path.join(os.path.dirname(__file__), "golden", "multi_index_ref.json")) as f:
        cases = json.load(f)["cases"]

    class FakeSegmented:
        tree = object()
        seg_trees = None

        def __init__(self, lists):
            self.lists = lists
            self.n_seg = len(lists)
            self.off = np.cumsum([0] + [len(x) for x in lists])

        def segment_of(self, o):
            s = int(np.searchsorted(self.off, o, side="right")) - 1
            return s, o - int(self.off[s])

        def retrieve_host(self, q, k, ratio, merge=True):
            w = max(1, max(len(x) for x in self.lists))
            ids = np.full((self.n_seg, 1, w), -1, np.int64)
            sc = np.zeros((self.n_seg, 1, w))
            lens = np.zeros((self.n_seg, 1), np.int32)
            for s, x in enumerate(self.lists):
                lens[s, 0] = len(x)
                for j, (_, score) in enumerate(x):
                    ids[s, 0, j] = self.off[s] + j
                    sc[s, 0, j] = score
            return ids, sc, lens

    ran = 0
    for case in cases:
        if any(x == "raise" for x in case["lists"]) or any(sc is None for x in case["lists"] for _, sc in x):
            continue  # a device index neither raises per segment nor reports a missing score
        idx = FakeSegmented(case["lists"])
        tables = [NodeTable(node_ids=[i for i, _ in x]) for x in case["lists"]]
        m = B200MultiIndexRetriever(idx, 10, embed_model=SimpleNamespace(get_agg_embedding_from_queries=lambda s: [0.0] * 8),
                                    node_tables=tables, balance_strategy=case["strategy"])
        got = [[n.node.id_, n.score, n.node.metadata["_source_index"]] for n in m.retrieve("the query")]
        if case["strategy"] == "none":
            got = sorted(got, key=lambda t: t[0])
        assert got == case["expected"], case["name"]
        ran += 1
    assert ran >= 15


def test_retriever_close_releases_the_index_and_the_cache():
    from tensor_truth_b200.retriever import B200AutoMergingRetriever, B200MultiIndexRetriever, B200VectorIndexRetriever

    closed = []
    idx = SimpleNamespace(tree=object(), close=lambda: closed.append("index"), n_seg=1, seg_trees=None)
    base = B200VectorIndexRetriever(idx, 10)
    B200AutoMergingRetriever(base, None).close()
    assert closed == ["index"]
    m = B200MultiIndexRetriever(idx, 10)
    m._retrieve_cached.__wrapped__  # an lru_cache wrapper
    m.close()
    assert closed == ["index", "index"] and m._retrieve_cached.cache_info().currsize == 0


def test_rerank_postprocessor_semantics_with_a_stand_in_encoder():
    """SentenceTransformerRerank.postprocess_nodes as restated in rerank.py: score overwrite with the sigmoid of the
    logit, sort descending, top_n cut, optional retrieval_score, the upstream error for a missing query bundle."""
    torch = pytest.importorskip("torch")
    from tensor_truth_b200.rerank import B200CrossEncoderRerank

    class Enc:
        max_length = 16

        def logits(self, toks):
            return torch.tensor([float(len(t)) - 5.0 for t in toks])

    seen = []

    def tokenize(pairs, max_length):
        seen.extend(pairs)
        return [[0] * min(max_length, len(d.split())) for _, d in pairs]

    nodes = [NodeWithScore(TextNode(id_=f"n{i}", text=" ".join(["w"] * (3 + 2 * i))), 0.5) for i in range(5)]
    rr = B200CrossEncoderRerank(Enc(), tokenize, top_n=2, keep_retrieval_score=True)
    out = rr.postprocess_nodes(list(nodes), query_bundle=QueryBundle(query_str="the question"))
    assert [n.node.id_ for n in out] == ["n4", "n3"]
    assert out[0].score == pytest.approx(1 / (1 + np.exp(-6.0))) and out[0].node.metadata["retrieval_score"] == 0.5
    assert all(q == "the question" for q, _ in seen) and len(seen) == 5
    assert rr.postprocess_nodes([], query_bundle=QueryBundle(query_str="q")) == []
    with pytest.raises(ValueError, match="Missing query bundle"):
        rr.postprocess_nodes(list(nodes))
    assert len(rr.postprocess_nodes(list(nodes), query_str="as a plain string")) == 2


# --------------------------------------------------------------------------- metadata filters (the filter tail)
_FILTER_SPECS = [
    {"doc_type": "library"},
    {"version": {"$gte": "2.0"}},
    {"doc_type": ["library", "book"]},
    {"doc_type": "paper", "year": {"$lt": 2020}, "lang": ["en", "de"]},
    {"x": {"$bogus": 1}},                      # unknown operator: the clause is dropped
    {"x": {"$ne": 3, "$gt": 1}},               # only the first key of an operator dict counts
    {"tags": {"$nin": ["draft"]}},
    {},
]


def test_filter_clauses_and_row_eligibility():
    from tensor_truth_b200.filters import clauses_from_spec, eligible_rows

    assert clauses_from_spec(_FILTER_SPECS[3]) == [("doc_type", "$eq", "paper"), ("year", "$lt", 2020), ("lang", "$in", ["en", "de"])]
    assert clauses_from_spec(_FILTER_SPECS[4]) == [] and clauses_from_spec(_FILTER_SPECS[5]) == [("x", "$ne", 3)]
    meta = [{"doc_type": "paper", "year": 2019, "lang": "en"}, {"doc_type": "paper", "year": 2021, "lang": "en"},
            {"doc_type": "book", "year": 2000, "lang": "de"}, {"doc_type": "paper", "lang": "de"}, None,
            {"doc_type": "paper", "year": "2019", "lang": "en"}]
    assert eligible_rows(_FILTER_SPECS[3], meta).tolist() == [True, False, False, False, False, False]  # missing key / wrong type: no match
    assert eligible_rows({"year": {"$ne": 2019}}, meta).tolist() == [False, True, True, True, True, True]  # $ne matches a missing key
    assert eligible_rows({"lang": {"$nin": ["en"]}}, meta).tolist() == [False, False, True, True, True, False]
    assert eligible_rows(None, meta).all() and eligible_rows({}, meta).all()
    with pytest.raises(ValueError):
        eligible_rows({"t": {"$text_match": "x"}}, meta)


def test_filter_spec_parser_equals_the_references_builder():
    """``clauses_from_spec`` against the reference's own ``_build_metadata_filters`` (rag_engine.py:301-365), executed as
    it lies under stand-ins for the LlamaIndex filter types (this container only)."""
    import importlib.util
    from dataclasses import dataclass, field
    from enum import Enum
    from typing import Any, List

    from tensor_truth_b200.filters import clauses_from_filters, clauses_from_spec

    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_multi_index_golden", os.path.join(here, "golden", "make_multi_index_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    if not gen.reference_available():
        pytest.skip("/root/reference is not on this machine")

    class FilterOperator(str, Enum):
        EQ, NE, GT, GTE, LT, LTE, IN, NIN, CONTAINS, TEXT_MATCH = "==", "!=", ">", ">=", "<", "<=", "in", "nin", "contains", "text_match"

    class FilterCondition(str, Enum):
        AND, OR = "and", "or"

    @dataclass
    class MetadataFilter:
        key: str
        value: Any
        operator: FilterOperator = FilterOperator.EQ

    @dataclass
    class MetadataFilters:
        filters: List[Any] = field(default_factory=list)
        condition: FilterCondition = FilterCondition.AND

    ref = gen.load_reference_rag_engine(dict(FilterOperator=FilterOperator, FilterCondition=FilterCondition,
                                             MetadataFilter=MetadataFilter, MetadataFilters=MetadataFilters))
    for fs in _FILTER_SPECS:
        built = ref._build_metadata_filters(fs)
        mine = clauses_from_spec(fs)
        assert (built is None) == (mine == []), fs
        if built is not None:
            assert clauses_from_filters(built) == mine, fs


# --------------------------------------------------------------------------- host lanes + coalescing (no GPU: a fake device)
def test_host_lanes_coalesce_concurrent_callers_without_starving_anyone(monkeypatch):
    """The lane / coalescing logic of ``DeviceIndex.retrieve_host`` against a fake device that takes 2 ms per pass
    whatever the batch: eight callers that re-enter immediately (no think time -- the pattern that starved the waiting
    callers of an earlier version) all get THEIR answers, batches form (so throughput exceeds one query per pass), batch
    shapes are powers of two, never more than HOST_LANES passes are in flight, and an error reaches the whole batch."""
    import contextlib
    import threading
    import time

    import torch

    from tensor_truth_b200 import index as im

    monkeypatch.setattr(torch.cuda, "Stream", lambda device=None: type("S", (), {"wait_stream": lambda self, o: None})())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda device=None: None)
    monkeypatch.setattr(torch.cuda, "stream", lambda st: contextlib.nullcontext())
    idx = object.__new__(im.DeviceIndex)
    idx._pending, idx._lane_cv, idx._coalesce, idx.coalesced = [], threading.Condition(), True, 0
    idx._free_lanes, idx._lane_streams = list(range(im.HOST_LANES)), {}
    idx.dim, idx.tree, idx.device = 4, None, None
    device, state, shapes = threading.Lock(), {"in_flight": 0, "peak": 0}, []

    def fake_pass(lane, q, k, ratio, merge, row_filter):
        if k < 0:
            raise ValueError("bad k")
        with idx._lane_cv:
            state["in_flight"] += 1
            state["peak"] = max(state["peak"], state["in_flight"])
            shapes.append(int(q.shape[0]))
        with device:
            time.sleep(0.002)
        with idx._lane_cv:
            state["in_flight"] -= 1
        tag = q[:, 0].numpy().astype(np.int64)  # the answer of a query is its first coordinate
        return tag[:, None].repeat(3, axis=1), tag[:, None].astype(np.float64), np.full(len(tag), 3, np.int32)

    idx._retrieve_host_lane = fake_pass
    wrong, per = [], 25

    def caller(t):
        for i in range(per):
            v = 1000 * t + i
            ids, scores, lens = idx.retrieve_host(torch.full((1, 4), float(v)), 10)
            if ids.shape != (1, 3) or int(ids[0, 0]) != v or float(scores[0, 0]) != v or int(lens[0]) != 3:
                wrong.append((t, i, ids))

    threads = [threading.Thread(target=caller, args=(t,)) for t in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=60)
    assert not any(th.is_alive() for th in threads), "a caller is stuck"
    assert not wrong and not idx._pending and sorted(idx._free_lanes) == list(range(im.HOST_LANES))
    assert state["peak"] <= im.HOST_LANES and idx.coalesced > 0
    assert all(b & (b - 1) == 0 and b <= im.COALESCE_MAX for b in shapes) and max(shapes) > 1
    assert len(shapes) < 8 * per  # fewer passes than queries: coalescing paid off
    # one caller alone leads a batch of itself
    n_before = idx.coalesced
    assert int(idx.retrieve_host(torch.full((1, 4), 7.0), 10)[0][0, 0]) == 7 and idx.coalesced == n_before
    # an error inside a coalesced batch reaches every caller of the batch
    held = idx._hold_all_lanes()
    errors = []

    def failing():
        try:
            idx.retrieve_host(torch.zeros((1, 4)), -1)
        except ValueError as exc:
            errors.append(exc)

    threads = [threading.Thread(target=failing) for _ in range(3)]
    for th in threads:
        th.start()
    t0 = time.time()
    while len(idx._pending) < 3 and time.time() - t0 < 10:
        time.sleep(0.002)
    idx._release_lanes(held)
    for th in threads:
        th.join(timeout=30)
    assert len(errors) == 3 and len({id(e) for e in errors}) == 1
