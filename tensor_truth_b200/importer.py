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
"""

from __future__ import annotations

from typing import Any, Dict, List, Optional, Sequence, Tuple

import numpy as np

from .tree import NodeTree, tree_from_relations


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
    Leaves the vector store knows but the docstore does not (or the reverse) are an error, as they would make
    ordinals and corpus rows disagree."""
    leaf_ids = list(leaf_ids)
    missing = [i for i in leaf_ids if i not in docs]
    if missing:
        raise ValueError(f"{len(missing)} embedded node ids are not in the docstore (first: {missing[0]!r})")
    parent, children, prev, nxt = relations_from_docstore(docs)
    embedded = set(leaf_ids)
    stray = [i for i, n in docs.items() if not children[i] and i not in embedded]
    if stray:
        raise ValueError(f"{len(stray)} leaf nodes of the docstore have no embedding (first: {stray[0]!r})")
    tree = tree_from_relations(list(docs), parent, children, prev, nxt, leaf_ids)
    corpus = np.ascontiguousarray(np.asarray(embeddings, dtype=np.float32))
    if corpus.ndim != 2 or corpus.shape[0] != len(leaf_ids):
        raise ValueError(f"embeddings shape {corpus.shape} does not match {len(leaf_ids)} ids")
    nodes = [docs[i] for i in tree.node_ids]
    return corpus, tree, nodes


def load_device_index(collection: Any, docstore: Any, device=None, **index_kw):
    """Snapshot a loaded reference index into HBM.  Returns ``(DeviceIndex, nodes_by_ordinal)``.
    The stored embeddings are fp32: they become the fp32 master, scanned through a bf16 shadow (index.py)."""
    from .index import DeviceIndex

    got = collection.get(include=["embeddings"])
    docs = docstore.docs if hasattr(docstore, "docs") else dict(docstore)
    corpus, tree, nodes = flatten_index(got["ids"], got["embeddings"], docs)
    return DeviceIndex(corpus, tree, device=device, **index_kw), nodes
