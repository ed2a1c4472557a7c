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
import copy
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

    def close(self) -> None:
        """Drop the device memory behind this retriever (corpus, tree arrays, workspaces, captured graphs).  The
        reference releases its retrievers in ``RAGService.clear()`` (rag_service.py:720-740: ``clear_cache()``, then the
        engine is dropped and ``torch.cuda.empty_cache()`` runs); deleting the retriever has the same effect here, and a
        patched ``clear()`` can call this to free HBM at once."""
        idx = getattr(self, "index", None)
        if idx is not None and hasattr(idx, "close"):
            idx.close()
        if hasattr(self, "clear_cache"):
            self.clear_cache()

    async def aretrieve(self, str_or_query_bundle: QueryType) -> List[NodeWithScore]:
        return self.retrieve(str_or_query_bundle)


class B200VectorIndexRetriever(_RetrieverBase):
    """``index.as_retriever(similarity_top_k=k)``: exact top-k leaves of the index for the query."""

    def __init__(self, index, similarity_top_k: int = 10, embed_model: Any = None,
                 node_table: Optional[NodeTable] = None, filters: Any = None,
                 leaf_metadata: Optional[Sequence[Optional[dict]]] = None, lane: Optional[int] = None):
        """``index``: a ``DeviceIndex``, or a ``ShardedIndex`` (row-sharded over GPUs; every rank then makes the same calls).
        ``filters``: what ``index.as_retriever(..., filters=...)`` takes upstream -- a LlamaIndex ``MetadataFilters`` or the
        reference's filter-spec dict (``_build_metadata_filters``, rag_engine.py:301-365); evaluated once, here, over
        ``leaf_metadata`` (one dict per corpus row; default: the ``.metadata`` of the node table's leaves) and applied
        to every search as a row gate (filters.py).  Single-GPU indexes only.
        ``lane``: for a ``ShardedIndex`` served by several threads -- thread t of every rank uses the retriever
        ``for_lane(t)`` (``ShardedIndex.retrieve_host``).  A ``DeviceIndex`` pipelines concurrent callers by itself."""
        self.index = index
        self.lane = lane
        self.similarity_top_k = int(similarity_top_k)
        self.embed_model = embed_model
        self.node_table = node_table or NodeTable(node_ids=getattr(index.tree, "node_ids", None) if index.tree else None)
        self.row_filter = None
        if filters is not None:
            from .filters import clauses_from_filters, eligible_rows

            clauses = clauses_from_filters(filters)
            if clauses:
                if not hasattr(index, "row_filter"):
                    raise ValueError("metadata filters need a single-GPU DeviceIndex")
                if leaf_metadata is None:
                    leaf_metadata = [getattr(self.node_table(o), "metadata", None) for o in range(index.n_rows)]
                self.row_filter = index.row_filter(eligible_rows(filters, leaf_metadata), key=repr(sorted(map(repr, clauses))))

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

    def _call_kw(self) -> dict:
        kw = {"row_filter": self.row_filter} if self.row_filter is not None else {}
        if self.lane is not None:
            kw["lane"] = self.lane
        return kw

    def for_lane(self, lane: int) -> "B200VectorIndexRetriever":
        """The same retriever bound to a host lane of a ``ShardedIndex`` (a shallow copy: index and node table shared)."""
        other = copy.copy(self)
        other.lane = int(lane)
        return other

    def _retrieve(self, query_bundle: QueryBundle) -> List[NodeWithScore]:
        kw = self._call_kw()
        ids, scores, lens = self.index.retrieve_host(self._query_tensor(query_bundle), self.similarity_top_k, merge=False, **kw)
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

    def for_lane(self, lane: int) -> "B200AutoMergingRetriever":
        """This retriever bound to a host lane of a ``ShardedIndex``: one per serving thread, the same lane on every rank."""
        other = copy.copy(self)
        other._vector_retriever = self._vector_retriever.for_lane(lane)
        return other

    def _retrieve(self, query_bundle: QueryBundle) -> List[NodeWithScore]:
        q = self._vector_retriever._query_tensor(query_bundle)
        kw = self._vector_retriever._call_kw()
        ids, scores, lens = self.index.retrieve_host(q, self._vector_retriever.similarity_top_k,
                                                     self._simple_ratio_thresh, merge=True, **kw)
        if lens[0] < 0:
            raise RuntimeError("auto-merge output overflow")
        return self._wrap(ids[0], scores[0], int(lens[0]))

    # ---- batch extension (not in the reference; what the bench drives)
    def retrieve_batch(self, embeddings) -> List[List[NodeWithScore]]:
        q = torch.as_tensor(np.asarray(embeddings, dtype=np.float32)) if not torch.is_tensor(embeddings) else embeddings
        kw = self._vector_retriever._call_kw()
        ids, scores, lens = self.index.retrieve_host(q, self._vector_retriever.similarity_top_k,
                                                     self._simple_ratio_thresh, merge=True, **kw)
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


class B200MultiIndexRetriever(_RetrieverBase):
    """``MultiIndexRetriever(retrievers=[AutoMergingRetriever(...) per index])`` (rag_engine.py:368-526, built at
    :688-692) over ONE ``SegmentedIndex`` holding all the indexes: the query is embedded once and one pass of the
    device pipeline answers every index, instead of m vector-store queries on a thread pool (:416-461).

    What the reference class does to the per-index results is kept as it is: every node is tagged
    ``metadata["_source_index"] = idx`` (:432-450); with more than one index and ``balance_strategy ==
    "top_k_per_index"`` each index keeps its first ``max(1, total // n_indexes)`` nodes and the union is sorted by
    ``score or 0.0``, descending, stable (:463-507); results are cached per query *string* in an LRU (:399-404) that
    ``clear_cache()`` empties (:520-526).  A segment that returns nothing simply contributes nothing (:453-455 is the
    reference's "skip a failing index")."""

    def __init__(self, index, similarity_top_k: int = 10, embed_model: Any = None,
                 node_tables: Optional[Sequence[NodeTable]] = None, simple_ratio_thresh: float = 0.5,
                 enable_cache: bool = True, cache_size: int = 128, balance_strategy: str = "top_k_per_index",
                 auto_merge: bool = True):
        from functools import lru_cache

        self.index = index
        self.similarity_top_k = int(similarity_top_k)
        self.embed_model = embed_model
        self.simple_ratio_thresh = float(simple_ratio_thresh)
        self.enable_cache = enable_cache
        self.balance_strategy = balance_strategy
        self.auto_merge = bool(auto_merge and index.tree is not None)
        if node_tables is None:
            trees = index.seg_trees or [None] * index.n_seg
            node_tables = [NodeTable(node_ids=getattr(t, "node_ids", None) if t is not None else None) for t in trees]
        if len(node_tables) != index.n_seg:
            raise ValueError("one node table per segment")
        self.node_tables = list(node_tables)
        self._embedder = B200VectorIndexRetriever.__new__(B200VectorIndexRetriever)  # reuse its bundle -> tensor logic
        self._embedder.embed_model = embed_model
        self._pending: dict = {}
        self._retrieve_cached = lru_cache(maxsize=cache_size)(self._retrieve_impl) if enable_cache else self._retrieve_impl

    def _retrieve_impl(self, query_text: str) -> List[NodeWithScore]:
        qb = self._pending.pop((threading.get_ident(), query_text), None) or QueryBundle(query_str=query_text)
        q = self._embedder._query_tensor(qb)
        ids, scores, lens = self.index.retrieve_host(q, self.similarity_top_k, self.simple_ratio_thresh, merge=self.auto_merge)
        combined: List[NodeWithScore] = []
        for s in range(self.index.n_seg):
            n = int(lens[s, 0])
            if n < 0:
                raise RuntimeError("auto-merge output overflow")
            for o, sc in zip(ids[s, 0, :n], scores[s, 0, :n]):
                seg, local = self.index.segment_of(int(o))
                node = self.node_tables[seg](local)
                if isinstance(getattr(node, "metadata", None), dict):
                    node.metadata["_source_index"] = s
                combined.append(NodeWithScore(node=node, score=float(sc)))
        if self.index.n_seg > 1 and self.balance_strategy == "top_k_per_index":
            combined = self._balance_top_k_per_index(combined)
        return combined

    @staticmethod
    def _balance_top_k_per_index(nodes: List[NodeWithScore]) -> List[NodeWithScore]:
        by_index: dict = {}
        for n in nodes:
            by_index.setdefault(n.node.metadata.get("_source_index", 0), []).append(n)
        if not by_index:
            return []
        limit = max(1, len(nodes) // len(by_index))
        kept = [n for group in by_index.values() for n in group[:limit]]
        kept.sort(key=lambda n: n.score if n.score else 0.0, reverse=True)
        return kept

    def _retrieve(self, query_bundle: QueryBundle) -> List[NodeWithScore]:
        # the cache key is the query string, like the reference's; a bundle that already carries an embedding hands it
        # to the (possibly cached-away) implementation through a side slot instead of being re-embedded
        if getattr(query_bundle, "embedding", None) is not None:
            self._pending[(threading.get_ident(), query_bundle.query_str)] = query_bundle
        try:
            return self._retrieve_cached(query_bundle.query_str)
        finally:
            self._pending.pop((threading.get_ident(), query_bundle.query_str), None)

    def clear_cache(self) -> None:
        if self.enable_cache and hasattr(self._retrieve_cached, "cache_clear"):
            self._retrieve_cached.cache_clear()
