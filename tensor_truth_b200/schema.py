"""The carrier types of the retriever protocol.

When ``llama_index`` is importable (any deployment of the reference) its own ``QueryBundle`` /
``NodeWithScore`` / ``TextNode`` are used, so objects flow into the rest of tensortruth unchanged.
Without it (this build image) the duck-typed stand-ins below expose exactly the attributes the
reference's consumers read:

* ``core/source_converter.py:84-137``      -- ``.node`` / ``.score`` / ``inner.id_`` / ``inner.metadata`` /
  ``inner.get_content()`` / ``inner.text``
* ``services/retrieval_metrics.py:163-247`` -- ``.score``, ``.node.metadata``, ``.node.get_content()``
* ``rag_engine.py:432-450``                 -- writes ``node.metadata["_source_index"]`` (needs a mutable dict)
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

try:  # pragma: no cover - llama_index is not installed in the build image
    from llama_index.core.schema import NodeWithScore, QueryBundle, TextNode  # type: ignore

    HAVE_LLAMA_INDEX = True
except Exception:  # ModuleNotFoundError, or a broken install
    HAVE_LLAMA_INDEX = False

    @dataclass
    class QueryBundle:  # type: ignore[no-redef]
        query_str: str
        embedding: Optional[List[float]] = None
        custom_embedding_strs: Optional[List[str]] = None

        @property
        def embedding_strs(self) -> List[str]:
            return self.custom_embedding_strs if self.custom_embedding_strs else [self.query_str]

    @dataclass
    class TextNode:  # type: ignore[no-redef]
        id_: str = ""
        text: str = ""
        metadata: Dict[str, Any] = field(default_factory=dict)
        parent_id: Optional[str] = None
        prev_id: Optional[str] = None
        next_id: Optional[str] = None
        child_ids: List[str] = field(default_factory=list)

        @property
        def node_id(self) -> str:
            return self.id_

        def get_content(self, metadata_mode: Any = None) -> str:
            return self.text

        # relationship accessors in upstream's spelling (RelatedNodeInfo-like: only node_id is read here)
        @property
        def parent_node(self):
            return _Related(self.parent_id) if self.parent_id is not None else None

        @property
        def prev_node(self):
            return _Related(self.prev_id) if self.prev_id is not None else None

        @property
        def next_node(self):
            return _Related(self.next_id) if self.next_id is not None else None

        @property
        def child_nodes(self):
            return [_Related(c) for c in self.child_ids] or None

    @dataclass
    class _Related:
        node_id: str

    @dataclass
    class NodeWithScore:  # type: ignore[no-redef]
        node: Any
        score: Optional[float] = None

        def get_score(self, raise_error: bool = False) -> float:
            if self.score is None:
                if raise_error:
                    raise ValueError("Score not set.")
                return 0.0
            return self.score

        # pass-throughs upstream's NodeWithScore also offers
        @property
        def node_id(self) -> str:
            return self.node.node_id

        @property
        def id_(self) -> str:
            return self.node.id_

        @property
        def text(self) -> str:
            return self.node.text

        @property
        def metadata(self) -> Dict[str, Any]:
            return self.node.metadata

        def get_content(self, metadata_mode: Any = None) -> str:
            return self.node.get_content(metadata_mode)
