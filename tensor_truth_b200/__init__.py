"""B200-native retrieval hot path for tensor-truth: dense scan -> exact top-k -> auto-merge.

Drop-in for the two objects the reference builds per index at
/root/reference/src/tensortruth/rag_engine.py:639-645 (``index.as_retriever`` +
``AutoMergingRetriever``).  All numeric work runs in hand-written sm_100a CUDA kernels
behind the C ABI declared in ``include/tt_b200.h``; there is no CPU fallback.
"""

__version__ = "0.1.0"

from .tree import NodeTree, build_uniform_tree, tree_from_relations  # noqa: F401
from .schema import NodeWithScore, QueryBundle, TextNode  # noqa: F401


def __getattr__(name):  # torch-dependent parts load lazily
    if name in ("DeviceIndex", "SearchResult", "MergeResult"):
        from . import index

        return getattr(index, name)
    if name == "SegmentedIndex":
        from . import segmented

        return segmented.SegmentedIndex
    if name in ("B200VectorIndexRetriever", "B200AutoMergingRetriever", "B200MultiIndexRetriever", "NodeTable", "build_retriever"):
        from . import retriever

        return getattr(retriever, name)
    raise AttributeError(name)
