"""Importer: the reference's on-disk index -> device arrays.

The reference persists an index as a Chroma collection ``"data"`` holding the LEAF embeddings
(/root/reference/src/tensortruth/indexing/builder.py:420-442) next to ``docstore.json`` holding ALL
nodes with their PARENT / CHILD / PREVIOUS / NEXT relations (builder.py:430,444).  At engine load
(rag_engine.py:628-636) both are open already; this module snapshots them into a ``DeviceIndex``:

    dev_index, nodes = load_device_index(collection, index.storage_context.docstore)

``collection`` only needs ``get(include=["embeddings"]) -> {"ids": [...], "embeddings": [...]}`` and the
docstore only ``docs: Dict[id, node]`` with upstream's relation accessors (``parent_node``, ``prev_node``,
``next_node``, ``child_nodes`` -> objects carrying ``node_id``), so the function is duck-typed and does not
import chromadb or llama_index itself.

Without ``llama_index`` (an export job, a GPU box that only serves) the relations can be read straight from the
persisted file: ``load_docstore_json(index_dir / "docstore.json")`` returns the same ``Dict[id, node]`` of light-weight
``StoredNode`` objects (``document_index.py:135-139`` names the file; the JSON layout is upstream's
``SimpleDocumentStore`` persistence **[U: restated from the public llama-index-core sources, no install to verify
against here]**: ``{"docstore/data": {id: {"__data__": node dict, "__type__": "1"}}, ...}`` with
``relationships`` keyed by ``NodeRelationship`` value -- "1" SOURCE, "2" PREVIOUS, "3" NEXT, "4" PARENT, "5" CHILD).
"""

from __future__ import annotations

import json
import os
from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .tree import NodeTree, tree_from_relations


# NodeRelationship values as persisted (upstream enum, [U]); names accepted too in case a writer used them
_REL_KEYS = {"source": ("1", "SOURCE"), "prev": ("2", "PREVIOUS"), "next": ("3", "NEXT"), "parent": ("4", "PARENT"),
             "child": ("5", "CHILD")}


class StoredRelation:
    """What upstream's ``RelatedNodeInfo`` carries, as far as this path reads it."""

    __slots__ = ("node_id", "node_type", "metadata", "hash")

    def __init__(self, d: Dict[str, Any]):
        self.node_id = d.get("node_id")
        self.node_type = d.get("node_type")
        self.metadata = d.get("metadata") or {}
        self.hash = d.get("hash")

    def __repr__(self) -> str:
        return f"StoredRelation({self.node_id!r})"


class StoredNode:
    """A docstore node read from ``docstore.json`` without llama_index: id, text, metadata and the four relations the
    auto-merge needs, behind upstream's accessor names (``TextNode.parent_node`` / ``prev_node`` / ``next_node`` /
    ``child_nodes``), so ``flatten_index`` and the retrievers' node table take it as they take the real class."""

    def __init__(self, data: Dict[str, Any], object_type: Optional[str] = None):
        self.id_ = data.get("id_") or data.get("node_id") or data.get("doc_id")
        self.text = data.get("text") or ""
        self.metadata: Dict[str, Any] = dict(data.get("metadata") or data.get("extra_info") or {})
        self.object_type = object_type
        self.class_name = data.get("class_name")
        self.start_char_idx = data.get("start_char_idx")
        self.end_char_idx = data.get("end_char_idx")
        rel = data.get("relationships") or {}

        def one(kind: str) -> Optional[StoredRelation]:
            for key in _REL_KEYS[kind]:
                v = rel.get(key)
                if isinstance(v, list):  # tolerated: a single relation persisted as a one-element list
                    v = v[0] if v else None
                if v:
                    return StoredRelation(v)
            return None

        self.source_node = one("source")
        self.prev_node = one("prev")
        self.next_node = one("next")
        self.parent_node = one("parent")
        kids = None
        for key in _REL_KEYS["child"]:
            if rel.get(key):
                kids = rel[key]
                break
        if isinstance(kids, dict):
            kids = [kids]
        self.child_nodes: Optional[List[StoredRelation]] = [StoredRelation(k) for k in kids] if kids else None

    @property
    def node_id(self) -> str:
        return self.id_

    def get_content(self, metadata_mode: Any = None) -> str:
        return self.text

    def get_text(self) -> str:
        return self.text

    def __repr__(self) -> str:
        return f"StoredNode({self.id_!r}, {len(self.text)} chars)"


def load_docstore_json(path: str) -> Dict[str, StoredNode]:
    """``docstore.json`` (or the index directory that holds it) -> ``{node id: StoredNode}`` in file order."""
    if os.path.isdir(path):
        path = os.path.join(path, "docstore.json")
    with open(path, "r", encoding="utf-8") as f:
        blob = json.load(f)
    data = blob.get("docstore/data")
    if data is None:
        raise ValueError(f"{path}: no 'docstore/data' collection (keys: {sorted(blob)[:5]}) -- not a SimpleDocumentStore file")
    out: Dict[str, StoredNode] = {}
    for key, wrapped in data.items():
        payload = wrapped.get("__data__", wrapped) if isinstance(wrapped, dict) else wrapped
        if isinstance(payload, str):  # some versions persist the node dict as a JSON string
            payload = json.loads(payload)
        node = StoredNode(payload, wrapped.get("__type__") if isinstance(wrapped, dict) else None)
        if node.id_ is None:
            node.id_ = key
        out[node.id_] = node
    return out


def _rel_id(node: Any, attr: str) -> Optional[str]:
    rel = getattr(node, attr, None)
    return None if rel is None else getattr(rel, "node_id", None)


def relations_from_docstore(docs: Dict[str, Any]):
    """``(parent, children, prev, next)`` id maps out of docstore nodes."""
    parent = {i: _rel_id(n, "parent_node") for i, n in docs.items()}
    prev = {i: _rel_id(n, "prev_node") for i, n in docs.items()}
    nxt = {i: _rel_id(n, "next_node") for i, n in docs.items()}
    children = {i: [getattr(c, "node_id", None) for c in (getattr(n, "child_nodes", None) or [])] for i, n in docs.items()}
    return parent, children, prev, nxt


def flatten_index(leaf_ids: Sequence[str], embeddings, docs: Dict[str, Any]) -> Tuple[np.ndarray, NodeTree, List[Any]]:
    """Host-side half of the import: ``(corpus fp32 [N, D] in leaf order, NodeTree, node objects by ordinal)``.
    An embedded id the docstore does not know is an error (the auto-merge could not place it).  A childless docstore
    node WITHOUT an embedding is kept and warned about: the reference would still serve such an index -- the node can
    never be retrieved, but it keeps counting in its parent's ``len(child_nodes)`` -- so it gets an ordinal after the
    leaves (no corpus row) and the engine load goes through."""
    leaf_ids = list(leaf_ids)
    missing = [i for i in leaf_ids if i not in docs]
    if missing:
        raise ValueError(f"{len(missing)} embedded node ids are not in the docstore (first: {missing[0]!r})")
    parent, children, prev, nxt = relations_from_docstore(docs)
    embedded = set(leaf_ids)
    stray = [i for i, n in docs.items() if not children[i] and i not in embedded]
    if stray:
        import warnings

        warnings.warn(f"{len(stray)} childless docstore nodes have no embedding and cannot be retrieved (first: {stray[0]!r})")
    tree = tree_from_relations(list(docs), parent, children, prev, nxt, leaf_ids)
    corpus = np.ascontiguousarray(np.asarray(embeddings, dtype=np.float32))
    if corpus.ndim != 2 or corpus.shape[0] != len(leaf_ids):
        raise ValueError(f"embeddings shape {corpus.shape} does not match {len(leaf_ids)} ids")
    nodes = [docs[i] for i in tree.node_ids]
    return corpus, tree, nodes


def collection_score_mode(collection: Any) -> int:
    """The score the reference surfaces for this collection.  It opens Chroma with ``get_or_create_collection("data")``
    (rag_engine.py:628-630, builder.py:424-426) and never sets ``hnsw:space``, so the space is Chroma's default,
    squared L2, and ``ChromaVectorStore`` reports ``exp(-distance)`` -> ``SCORE_CHROMA_L2_EXP``.  The auto-merge averages
    children's scores, and a mean is not invariant under that transform: only this mode reproduces the reference's
    merged-parent scores (and with them its order and the per-index truncation of ``_balance_top_k_per_index``).
    A collection created with another space is refused rather than silently scored differently."""
    from ._lib import SCORE_CHROMA_L2_EXP

    meta = getattr(collection, "metadata", None) or {}
    space = meta.get("hnsw:space", "l2") if isinstance(meta, dict) else "l2"
    cfg = getattr(collection, "configuration_json", None)
    if isinstance(cfg, dict):
        space = (cfg.get("hnsw") or {}).get("space", space) or space
    if space != "l2":
        raise ValueError(f"collection space {space!r}: only Chroma's default squared-L2 space is reproduced; "
                         "pass score_mode explicitly to override")
    return SCORE_CHROMA_L2_EXP


def load_device_index(collection: Any, docstore: Any, device=None, score_mode: Optional[int] = None, **index_kw):
    """Snapshot a loaded reference index into HBM.  Returns ``(DeviceIndex, nodes_by_ordinal)``.
    ``docstore``: the loaded docstore (``index.storage_context.docstore``), a ``{id: node}`` dict, or the path of the
    persisted ``docstore.json`` / its index directory.
    The stored embeddings are fp32: they become the fp32 master, scanned through a bf16 shadow (index.py).
    ``score_mode`` defaults to what the reference reports for the collection (``collection_score_mode``:
    ``exp(-squared L2)``); ``SCORE_COSINE`` is the explicit choice of the synthetic benchmark."""
    from .index import DeviceIndex

    if score_mode is None:
        score_mode = collection_score_mode(collection)
    index_kw["score_mode"] = score_mode
    got = collection.get(include=["embeddings"])
    if isinstance(docstore, (str, os.PathLike)):  # the persisted file (or the index directory): no llama_index needed
        docstore = load_docstore_json(os.fspath(docstore))
    docs = docstore.docs if hasattr(docstore, "docs") else dict(docstore)
    corpus, tree, nodes = flatten_index(got["ids"], got["embeddings"], docs)
    return DeviceIndex(corpus, tree, device=device, **index_kw), nodes
