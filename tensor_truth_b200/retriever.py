"""The retriever surface the rest of tensortruth calls, backed by the B200 kernels.

Drop-in for the two objects the reference builds per index at
/root/reference/src/tensortruth/rag_engine.py:639-645 (modules) and :674-679 (session/project
indexes)::

    base_retriever = index.as_retriever(similarity_top_k=similarity_top_k)
    am_retriever   = AutoMergingRetriever(base_retriever, index.storage_context, verbose=False)

``B200VectorIndexRetriever`` stands for the first, ``B200AutoMergingRetriever`` for the second.  Both
follow the LlamaIndex ``BaseRetriever`` protocol the callers rely on: ``retrieve(str | QueryBundle)``
(rag_service.py:320,594 pass a ``str``; ``MultiIndexRetriever`` passes a ``QueryBundle`` from worker
threads, rag_engine.py:416-424) and ``_retrieve(QueryBundle)``; the result is a
``List[NodeWithScore]`` sorted by score, descending.  Errors surface as exceptions
(``MultiIndexRetriever`` logs and skips a failing child, rag_engine.py:453-455); there is no CPU
fallback.
"""

from __future__ import annotations

import contextlib
import threading
from array import array
from typing import Any, Callable, List, Optional, Sequence, Union

import numpy as np
import torch

from ._lib import SCORE_COSINE
from .index import DeviceIndex
from .schema import NodeWithScore, QueryBundle, TextNode

QueryType = Union[str, QueryBundle]


class NodeTable:
    """ordinal -> node object.  Built by the importer from the docstore (tree.tree_from_relations keeps
    the ordinal -> node-id order); synthetic indexes create placeholder ``TextNode``s on demand."""

    def __init__(self, nodes: Optional[Sequence[Any]] = None, node_ids: Optional[Sequence[str]] = None,
                 factory: Optional[Callable[[int], Any]] = None):
        self._nodes = nodes
        self._ids = node_ids
        self._factory = factory
        self._made: dict = {}  # like the docstore, hand out the SAME node object for an ordinal every time

    def __call__(self, ordinal: int):
        if self._nodes is not None:
            return self._nodes[ordinal]
        node = self._made.get(ordinal)
        if node is None:
            if self._factory is not None:
                node = self._factory(ordinal)
            else:
                nid = self._ids[ordinal] if self._ids is not None else f"node-{ordinal}"
                node = TextNode(id_=nid, text="", metadata={})
            if len(self._made) < 1_000_000:
                self._made[ordinal] = node
        return node


def _embed(embed_model, query_bundle: QueryBundle) -> List[float]:
    """What ``VectorIndexRetriever._retrieve`` does when the bundle carries no embedding."""
    if embed_model is None:
        raise ValueError("query has no embedding and the retriever has no embed_model")
    strs = getattr(query_bundle, "embedding_strs", None) or [query_bundle.query_str]
    if hasattr(embed_model, "get_agg_embedding_from_queries"):
        return embed_model.get_agg_embedding_from_queries(strs)
    return embed_model.get_query_embedding(strs[0])


class _RetrieverBase:
    def retrieve(self, str_or_query_bundle: QueryType) -> List[NodeWithScore]:
        qb = QueryBundle(str_or_query_bundle) if isinstance(str_or_query_bundle, str) else str_or_query_bundle
        return self._retrieve(qb)

    async def aretrieve(self, str_or_query_bundle: QueryType) -> List[NodeWithScore]:
        return self.retrieve(str_or_query_bundle)


class B200VectorIndexRetriever(_RetrieverBase):
    """``index.as_retriever(similarity_top_k=k)``: exact top-k leaves of the index for the query."""

    def __init__(self, index, similarity_top_k: int = 10, embed_model: Any = None,
                 node_table: Optional[NodeTable] = None):
        """``index``: a ``DeviceIndex``, or a ``ShardedIndex`` (row-sharded over GPUs; every rank then makes the same calls)."""
        self.index = index
        self.similarity_top_k = int(similarity_top_k)
        self.embed_model = embed_model
        self.node_table = node_table or NodeTable(node_ids=getattr(index.tree, "node_ids", None) if index.tree else None)

    def _query_tensor(self, query_bundle: QueryBundle) -> torch.Tensor:
        emb = query_bundle.embedding
        if emb is None:
            # MultiIndexRetriever hands ONE bundle without an embedding to every per-index retriever on a thread pool
            # (rag_engine.py:416-424), and upstream embeds the same string once per index.  Embed it once: the
            # first thread does the work under a lock that lives on the bundle, the others pick up the result.
            try:
                lock = query_bundle.__dict__.setdefault("_tt_embed_lock", threading.Lock())
            except AttributeError:  # a bundle type without __dict__
                lock = contextlib.nullcontext()
            with lock:
                emb = query_bundle.embedding
                if emb is None:
                    emb = _embed(self.embed_model, query_bundle)
                    try:
                        query_bundle.embedding = emb  # upstream caches it on the bundle too
                    except Exception:
                        pass
        if isinstance(emb, (list, tuple)):  # what an embed model hands over; array('f') is the fastest list -> fp32 path
            return torch.frombuffer(array("f", emb), dtype=torch.float32).reshape(1, -1)
        return torch.as_tensor(np.asarray(emb, dtype=np.float32)).reshape(1, -1)

    def _retrieve(self, query_bundle: QueryBundle) -> List[NodeWithScore]:
        ids, scores, lens = self.index.retrieve_host(self._query_tensor(query_bundle), self.similarity_top_k, merge=False)
        return [NodeWithScore(node=self.node_table(int(o)), score=float(s))
                for o, s in zip(ids[0, :lens[0]], scores[0, :lens[0]])]


class B200AutoMergingRetriever(_RetrieverBase):
    """``AutoMergingRetriever(base_retriever, storage_context, simple_ratio_thresh=0.5, verbose=False)``.

    ``storage_context`` is accepted for signature compatibility; the relations it holds were flattened
    into ``index.tree`` at load time (tree.tree_from_relations)."""

    def __init__(self, vector_retriever: B200VectorIndexRetriever, storage_context: Any = None,
                 simple_ratio_thresh: float = 0.5, verbose: bool = False, **_ignored: Any):
        if vector_retriever.index.tree is None:
            raise ValueError("auto-merging needs the node tree (index.set_tree)")
        self._vector_retriever = vector_retriever
        self._storage_context = storage_context
        self._simple_ratio_thresh = float(simple_ratio_thresh)
        self._verbose = verbose
        self.index = vector_retriever.index
        self.node_table = vector_retriever.node_table

    def _wrap(self, ids, scores, n) -> List[NodeWithScore]:
        return [NodeWithScore(node=self.node_table(int(o)), score=float(s)) for o, s in zip(ids[:n], scores[:n])]

    def _retrieve(self, query_bundle: QueryBundle) -> List[NodeWithScore]:
        q = self._vector_retriever._query_tensor(query_bundle)
        ids, scores, lens = self.index.retrieve_host(q, self._vector_retriever.similarity_top_k,
                                                     self._simple_ratio_thresh, merge=True)
        if lens[0] < 0:
            raise RuntimeError("auto-merge output overflow")
        return self._wrap(ids[0], scores[0], int(lens[0]))

    # ---- batch extension (not in the reference; what the bench drives)
    def retrieve_batch(self, embeddings) -> List[List[NodeWithScore]]:
        q = torch.as_tensor(np.asarray(embeddings, dtype=np.float32)) if not torch.is_tensor(embeddings) else embeddings
        ids, scores, lens = self.index.retrieve_host(q, self._vector_retriever.similarity_top_k,
                                                     self._simple_ratio_thresh, merge=True)
        if (lens < 0).any():
            raise RuntimeError("auto-merge output overflow")
        return [self._wrap(ids[b], scores[b], int(lens[b])) for b in range(ids.shape[0])]


def build_retriever(corpus, tree, similarity_top_k: int = 10, embed_model: Any = None, nodes: Optional[Sequence[Any]] = None,
                    simple_ratio_thresh: float = 0.5, score_mode: int = SCORE_COSINE, **index_kw: Any) -> B200AutoMergingRetriever:
    """What a patched ``load_engine_for_modules`` calls instead of rag_engine.py:639-645."""
    index = DeviceIndex(corpus, tree, score_mode=score_mode, **index_kw)
    table = NodeTable(nodes=nodes) if nodes is not None else None
    base = B200VectorIndexRetriever(index, similarity_top_k, embed_model, table)
    return B200AutoMergingRetriever(base, None, simple_ratio_thresh)
