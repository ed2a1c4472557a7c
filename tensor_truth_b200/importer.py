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
    The stored embeddings are fp32: they become the fp32 master, scanned through a bf16 shadow (index.py).
    ``score_mode`` defaults to what the reference reports for the collection (``collection_score_mode``:
    ``exp(-squared L2)``); ``SCORE_COSINE`` is the explicit choice of the synthetic benchmark."""
    from .index import DeviceIndex

    if score_mode is None:
        score_mode = collection_score_mode(collection)
    index_kw["score_mode"] = score_mode
    got = collection.get(include=["embeddings"])
    docs = docstore.docs if hasattr(docstore, "docs") else dict(docstore)
    corpus, tree, nodes = flatten_index(got["ids"], got["embeddings"], docs)
    return DeviceIndex(corpus, tree, device=device, **index_kw), nodes
