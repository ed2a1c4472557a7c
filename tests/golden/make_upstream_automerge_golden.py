#!/usr/bin/env python
"""Pin the auto-merge rows to the REAL upstream implementation.

The build image has no ``llama-index-core`` (and no network), so the goldens under ``tests/golden/`` were produced
by this repo's own restatement (``oracle/automerge.py``) -- "parity unpinned".  Run this script on ANY box that has
``llama-index-core`` installed (``pip install llama-index-core``; no GPU, no tensor-truth needed):

    python tests/golden/make_upstream_automerge_golden.py

It drives ``llama_index.core.retrievers.AutoMergingRetriever`` -- the class the reference instantiates at
/root/reference/src/tensortruth/rag_engine.py:641-643, ``simple_ratio_thresh`` left at its default -- over

  (a) every hand-built case of ``tests/golden/automerge_handbuilt.json`` (relations turned into real ``TextNode``
      relationships in a real ``SimpleDocumentStore``; the "vector retriever" is a stub that returns the case's input
      list), and
  (b) trees produced by the real ``HierarchicalNodeParser`` over synthetic text (when a tokenizer is available
      offline), whose relations are exported so the CPU oracle can be run on exactly the same tree -- this also
      covers upstream comparing ``RelatedNodeInfo`` objects (id AND hash/metadata) where the oracle compares ids,

and writes ``tests/golden/upstream_automerge.json``.  ``tests/test_upstream_pin.py`` consumes that file when it is
present (oracle vs upstream: ids equal, float64 scores bit-equal) and skips with a "parity unpinned" message when
it is not.  Commit the JSON to turn rows A4-A6 from "partial" to pinned.
"""

from __future__ import annotations

import json
import os
import platform
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def _imports():
    from llama_index.core import StorageContext
    from llama_index.core.base.base_retriever import BaseRetriever
    from llama_index.core.retrievers import AutoMergingRetriever
    from llama_index.core.schema import NodeRelationship, NodeWithScore, QueryBundle, RelatedNodeInfo, TextNode
    from llama_index.core.storage.docstore import SimpleDocumentStore

    return (StorageContext, BaseRetriever, AutoMergingRetriever, NodeRelationship, NodeWithScore, QueryBundle,
            RelatedNodeInfo, TextNode, SimpleDocumentStore)


def run_case(case, mods):
    """One hand-built case through the real AutoMergingRetriever.  Returns ``[[ordinal, score], ...]``."""
    (StorageContext, BaseRetriever, AutoMergingRetriever, NodeRelationship, NodeWithScore, QueryBundle, RelatedNodeInfo,
     TextNode, SimpleDocumentStore) = mods
    n = len(case["parent_of"])
    nid = lambda o: f"node-{o:04d}"  # noqa: E731
    nodes = [TextNode(id_=nid(o), text=f"text of node {o}") for o in range(n)]
    kids = {o: [] for o in range(n)}
    for o in range(n):
        p = case["parent_of"][o]
        if p >= 0:
            kids[p].append(o)
    for o, node in enumerate(nodes):
        p = case["parent_of"][o]
        if p >= 0:
            node.relationships[NodeRelationship.PARENT] = RelatedNodeInfo(node_id=nid(p))
        # child_count is authoritative (a hand-built case may declare more children than appear in the arrays)
        declared = case["child_count"][o]
        child_ids = [nid(c) for c in kids[o]]
        child_ids += [f"ghost-{o}-{j}" for j in range(max(0, declared - len(child_ids)))]
        if child_ids:
            node.relationships[NodeRelationship.CHILD] = [RelatedNodeInfo(node_id=c) for c in child_ids[:max(declared, 0)] or child_ids]
        if case["prev_id"][o] >= 0:
            node.relationships[NodeRelationship.PREVIOUS] = RelatedNodeInfo(node_id=nid(case["prev_id"][o]))
        if case["next_id"][o] >= 0:
            node.relationships[NodeRelationship.NEXT] = RelatedNodeInfo(node_id=nid(case["next_id"][o]))
    docstore = SimpleDocumentStore()
    docstore.add_documents(nodes)
    ctx = StorageContext.from_defaults(docstore=docstore)
    initial = [(int(o), float(s)) for o, s in case["input"]]

    class Stub(BaseRetriever):
        def _retrieve(self, query_bundle):
            return [NodeWithScore(node=docstore.get_document(nid(o)), score=s) for o, s in initial]

    am = AutoMergingRetriever(Stub(), ctx, simple_ratio_thresh=float(case.get("ratio_thresh", 0.5)), verbose=False)
    out = am.retrieve(QueryBundle(query_str="q"))
    return [[int(x.node.node_id.split("-")[1]), float(x.score)] for x in out]


def parsed_tree_cases(mods, n_cases=24, seed=5):
    """Trees from the real HierarchicalNodeParser + random initial lists; exports the relation arrays."""
    (StorageContext, BaseRetriever, AutoMergingRetriever, NodeRelationship, NodeWithScore, QueryBundle, RelatedNodeInfo,
     TextNode, SimpleDocumentStore) = mods
    from llama_index.core import Document
    from llama_index.core.node_parser import HierarchicalNodeParser, get_leaf_nodes

    rng = random.Random(seed)
    words = ["alpha", "beta", "gamma", "delta", "tensor", "truth", "index", "query", "merge", "node", "leaf", "score"]
    docs = []
    for d in range(3):
        sents = [" ".join(rng.choice(words) for _ in range(rng.randint(6, 14))).capitalize() + "." for _ in range(400)]
        docs.append(Document(text=" ".join(sents), doc_id=f"doc-{d}"))
    parser = HierarchicalNodeParser.from_defaults(chunk_sizes=[512, 128, 32], chunk_overlap=4)
    nodes = parser.get_nodes_from_documents(docs)
    leaves = get_leaf_nodes(nodes)
    leaf_ids = [x.node_id for x in leaves]
    order = leaf_ids + [x.node_id for x in nodes if x.node_id not in set(leaf_ids)]
    ordinal = {i: o for o, i in enumerate(order)}
    by_id = {x.node_id: x for x in nodes}

    def rel(x, name):
        r = getattr(x, name)
        return ordinal.get(r.node_id, -1) if r is not None else -1

    tree = {"parent_of": [rel(by_id[i], "parent_node") for i in order],
            "child_count": [len(by_id[i].child_nodes or []) for i in order],
            "prev_id": [rel(by_id[i], "prev_node") for i in order],
            "next_id": [rel(by_id[i], "next_node") for i in order], "n_leaf": len(leaf_ids)}
    docstore = SimpleDocumentStore()
    docstore.add_documents(nodes)
    ctx = StorageContext.from_defaults(docstore=docstore)
    cases = []
    for c in range(n_cases):
        k = rng.choice([5, 10, 20, 40])
        start = rng.randrange(0, max(1, len(leaf_ids) - 3 * k))
        picks = sorted(rng.sample(range(start, min(len(leaf_ids), start + 3 * k)), min(k, len(leaf_ids) - start)))
        rng.shuffle(picks)
        scores = sorted((rng.random() for _ in picks), reverse=True)
        initial = list(zip(picks, scores))

        class Stub(BaseRetriever):
            def _retrieve(self, query_bundle, initial=initial):
                return [NodeWithScore(node=docstore.get_document(order[o]), score=s) for o, s in initial]

        out = AutoMergingRetriever(Stub(), ctx, verbose=False).retrieve(QueryBundle(query_str="q"))
        cases.append({"input": [[int(o), float(s)] for o, s in initial],
                      "got": [[ordinal[x.node.node_id], float(x.score)] for x in out]})
    return {"tree": tree, "cases": cases, "chunk_sizes": [512, 128, 32], "chunk_overlap": 4}


def main():
    try:
        mods = _imports()
    except ImportError as exc:
        sys.exit(f"llama-index-core is not importable here ({exc}); run this on a box that has it")
    import llama_index.core as lic

    with open(os.path.join(HERE, "automerge_handbuilt.json")) as f:
        hand = json.load(f)
    out = {"llama_index_core_version": getattr(lic, "__version__", "unknown"), "python": platform.python_version(),
           "handbuilt": [{"label": c["label"], "got": run_case(c, mods)} for c in hand]}
    try:
        out["parsed"] = parsed_tree_cases(mods)
    except Exception as exc:  # e.g. no tokenizer available offline
        out["parsed"] = None
        out["parsed_error"] = f"{type(exc).__name__}: {exc}"
    dst = os.path.join(HERE, "upstream_automerge.json")
    with open(dst, "w") as f:
        json.dump(out, f, indent=1)
    agree = sum(1 for c, g in zip(hand, out["handbuilt"]) if [[int(o), float(s)] for o, s in c["expected"]] == g["got"])
    print(f"wrote {dst}: {len(hand)} hand-built cases ({agree} agree with the committed oracle goldens), "
          f"parsed-tree cases: {len(out['parsed']['cases']) if out.get('parsed') else out.get('parsed_error')}")


if __name__ == "__main__":
    main()
