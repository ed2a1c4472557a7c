"""Oracle, part 3: the caller of the hot path -- a restatement of ``MultiIndexRetriever``
(/root/reference/src/tensortruth/rag_engine.py:368-526).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  The reference keeps this class as is when
the B200 retrievers are dropped in; it cannot be imported here (``rag_engine.py:7-8`` imports
``llama_index``), so the tests drive the new retrievers through this restatement instead:

* ``_retrieve`` -> LRU cache keyed by the query *string* (:399-404, :509-518);
* ``_retrieve_impl`` rebuilds ``QueryBundle(query_str=...)`` (no embedding: every child embeds the string
  itself), fans it out on a ``ThreadPoolExecutor(max_workers=min(n, 8))`` (:392, :416-424), tags
  ``node.metadata["_source_index"]`` (:432-450), skips a child that raises (:453-455);
* with more than one index and ``balance_strategy == "top_k_per_index"``: keep the first
  ``max(1, total // n_indexes)`` nodes of each index, then sort by ``score or 0.0`` descending (:463-507).
"""

from __future__ import annotations

from collections import defaultdict
from concurrent.futures import ThreadPoolExecutor, as_completed
from functools import lru_cache


class MultiIndexRetriever:
    def __init__(self, retrievers, max_workers=None, enable_cache=True, cache_size=128,
                 balance_strategy="top_k_per_index", query_bundle_cls=None):
        self.retrievers = retrievers
        self.max_workers = max_workers or min(len(retrievers), 8)
        self.enable_cache = enable_cache
        self.balance_strategy = balance_strategy
        if query_bundle_cls is None:
            from tensor_truth_b200.schema import QueryBundle as query_bundle_cls
        self._qb = query_bundle_cls
        self._retrieve_cached = lru_cache(maxsize=cache_size)(self._retrieve_impl) if enable_cache else self._retrieve_impl

    def _retrieve_impl(self, query_text):
        bundle = self._qb(query_str=query_text)
        combined = []
        with ThreadPoolExecutor(max_workers=self.max_workers) as pool:
            futures = {pool.submit(r.retrieve, bundle): i for i, r in enumerate(self.retrievers)}
            for fut in as_completed(futures):
                try:
                    nodes = fut.result()
                except Exception as exc:  # a failing index is reported and skipped
                    print(f"Retriever failed: {exc}")
                    continue
                i = futures[fut]
                for n in nodes:
                    inner = getattr(n, "node", None)
                    if inner is not None and isinstance(getattr(inner, "metadata", None), dict):
                        inner.metadata["_source_index"] = i
                    elif isinstance(getattr(n, "metadata", None), dict):
                        n.metadata["_source_index"] = i
                combined.extend(nodes)
        if len(self.retrievers) > 1 and self.balance_strategy == "top_k_per_index":
            combined = self._balance_top_k_per_index(combined)
        return combined

    @staticmethod
    def _balance_top_k_per_index(nodes):
        groups = defaultdict(list)
        for n in nodes:
            meta = None
            if hasattr(n, "node") and hasattr(n.node, "metadata"):
                meta = n.node.metadata
            elif hasattr(n, "metadata"):
                meta = n.metadata
            groups[meta.get("_source_index", 0) if meta else 0].append(n)
        if not groups:
            return []
        limit = max(1, len(nodes) // len(groups))
        kept = [n for g in groups.values() for n in g[:limit]]
        kept.sort(key=lambda n: n.score if n.score else 0.0, reverse=True)
        return kept

    def _retrieve(self, query_bundle):
        return self._retrieve_cached(query_bundle.query_str)

    def retrieve(self, str_or_query_bundle):
        qb = self._qb(query_str=str_or_query_bundle) if isinstance(str_or_query_bundle, str) else str_or_query_bundle
        return self._retrieve(qb)

    def clear_cache(self):
        if self.enable_cache and hasattr(self._retrieve_cached, "cache_clear"):
            self._retrieve_cached.cache_clear()
