"""B200-native retrieval hot path for tensor-truth: dense scan -> exact top-k -> auto-merge.

Drop-in for the two objects the reference builds per index at
/root/reference/src/tensortruth/rag_engine.py:639-645 (``index.as_retriever`` +
``AutoMergingRetriever``).  All numeric work runs in hand-written sm_100a CUDA kernels
behind the C ABI declared in ``include/tt_b200.h``; there is no CPU fallback.
"""

__version__ = "0.1.0"

from .tree import NodeTree, build_uniform_tree, tree_from_relations  # noqa: F401
