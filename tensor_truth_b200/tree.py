"""Flat-array encoding of the hierarchical node tree the auto-merge step walks.

The reference keeps these relations as ``TextNode.relationships`` inside
``docstore.json`` -- written by ``HierarchicalNodeParser`` + ``docstore.add_documents``
at /root/reference/src/tensortruth/indexing/builder.py:385-430 (all levels stored,
leaves only embedded, :420-442).  The device kernels need them as arrays indexed by
a node *ordinal*:

* ordinals ``[0, n_leaf)`` are the leaves, in corpus-row order (ordinal == row);
* then level L-2 (the leaves' parents), ..., finally level 0 (no parent);
* ``parent_of[o]`` (i32, -1 none), ``child_count[o]`` (i32, 0 for leaves),
  ``prev_id[o]`` / ``next_id[o]`` (i32, -1 none) -- PREVIOUS/NEXT exist only between
  consecutive nodes split from the same source (siblings; level-0 nodes of one document).
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np


@dataclass
class NodeTree:
    parent_of: np.ndarray
    child_count: np.ndarray
    prev_id: np.ndarray
    next_id: np.ndarray
    n_leaf: int
    level_offsets: List[int] = field(default_factory=list)  # ordinal where each level starts, leaves first
    node_ids: Optional[List[str]] = None  # ordinal -> docstore node id (host only)

    @property
    def n_nodes(self) -> int:
        return int(self.parent_of.shape[0])

    def validate(self) -> None:
        n = self.n_nodes
        for name in ("parent_of", "child_count", "prev_id", "next_id"):
            a = getattr(self, name)
            if a.dtype != np.int32 or a.shape != (n,):
                raise ValueError(f"{name}: expected int32[{n}], got {a.dtype}{a.shape}")
        if n and (self.parent_of.max() >= n or self.prev_id.max() >= n or self.next_id.max() >= n):
            raise ValueError("relation ordinal out of range")
        if self.n_leaf > n:
            raise ValueError("n_leaf > n_nodes")


def build_uniform_tree(n_leaf: int, levels: int = 3, seed: int = 1234,
                       fan_lo: int = 2, fan_hi: int = 6) -> NodeTree:
    """Synthetic hierarchy of SURVEY.md section 8d: ``levels`` node levels (3 = the
    2048/512/128-token default of ``HierarchicalNodeParser``), fan-out per internal node
    uniform{fan_lo..fan_hi}; one extra virtual level groups level-0 nodes into documents
    so they get PREVIOUS/NEXT links like upstream, but no PARENT."""
    if levels < 1:
        raise ValueError("levels >= 1")
    rng = np.random.default_rng(seed)
    sizes = [int(n_leaf)]
    parents_local = []  # per level (leaves first): index of the parent within the next level up
    for _ in range(levels):  # levels-1 real parent levels + 1 virtual document level
        n = sizes[-1]
        if n == 0:
            parents_local.append(np.zeros(0, dtype=np.int64))
            sizes.append(0)
            continue
        fan = rng.integers(fan_lo, fan_hi + 1, size=n // fan_lo + 1, dtype=np.int64)
        csum = np.cumsum(fan)
        m = int(np.searchsorted(csum, n, side="left")) + 1
        fan = fan[:m].copy()
        fan[-1] -= csum[m - 1] - n
        parents_local.append(np.repeat(np.arange(m, dtype=np.int64), fan))
        sizes.append(m)
    real_sizes = sizes[:levels]
    offsets = np.concatenate([[0], np.cumsum(real_sizes)]).astype(np.int64)
    n_nodes = int(offsets[-1])
    if n_nodes >= 2**31:
        raise ValueError("node count exceeds int32 ordinals")
    parent_of = np.full(n_nodes, -1, dtype=np.int32)
    child_count = np.zeros(n_nodes, dtype=np.int32)
    prev_id = np.full(n_nodes, -1, dtype=np.int32)
    next_id = np.full(n_nodes, -1, dtype=np.int32)
    for lv in range(levels):
        lo, hi = int(offsets[lv]), int(offsets[lv + 1])
        n = hi - lo
        if n == 0:
            continue
        pl = parents_local[lv]
        if lv + 1 < levels:
            parent_of[lo:hi] = (pl + offsets[lv + 1]).astype(np.int32)
            child_count[offsets[lv + 1]:offsets[lv + 2]] = np.bincount(pl, minlength=real_sizes[lv + 1]).astype(np.int32)
        same = pl[1:] == pl[:-1]
        ords = np.arange(lo, hi, dtype=np.int32)
        prev_id[lo + 1:hi] = np.where(same, ords[:-1], -1)
        next_id[lo:hi - 1] = np.where(same, ords[1:], -1)
    tree = NodeTree(parent_of, child_count, prev_id, next_id, int(n_leaf), [int(x) for x in offsets[:-1]])
    return tree


def tree_from_relations(node_ids, parent, children, prev, nxt, leaf_order) -> NodeTree:
    """Importer core: build the arrays from docstore-style relations.

    ``node_ids``: all node ids; ``parent/prev/nxt``: dict id -> id or None; ``children``: dict
    id -> list of ids; ``leaf_order``: leaf ids in corpus-row order (what was embedded into the
    vector store, builder.py:420-442)."""
    leaf_set = set(leaf_order)
    internal = [i for i in node_ids if i not in leaf_set]

    def depth(i):
        d = 0
        while parent.get(i) is not None:
            i = parent[i]
            d += 1
        return d

    internal.sort(key=lambda i: -depth(i))  # deeper (closer to leaves) first; stable within level
    order = list(leaf_order) + internal
    ordinal = {nid: o for o, nid in enumerate(order)}
    n = len(order)
    parent_of = np.full(n, -1, dtype=np.int32)
    child_count = np.zeros(n, dtype=np.int32)
    prev_id = np.full(n, -1, dtype=np.int32)
    next_id = np.full(n, -1, dtype=np.int32)
    for nid, o in ordinal.items():
        p = parent.get(nid)
        if p is not None and p in ordinal:
            parent_of[o] = ordinal[p]
        child_count[o] = len(children.get(nid) or [])
        a = prev.get(nid)
        if a is not None and a in ordinal:
            prev_id[o] = ordinal[a]
        b = nxt.get(nid)
        if b is not None and b in ordinal:
            next_id[o] = ordinal[b]
    return NodeTree(parent_of, child_count, prev_id, next_id, len(leaf_order), [], list(order))


def concat_trees(trees: List[NodeTree]):
    """Several indexes' trees as ONE ordinal space (multi-index as one segmented corpus).

    Combined ordinals: the leaves of all segments first, in segment order (ordinal == row of the concatenated
    corpus), then the internal nodes of segment 0, of segment 1, ...  Returns ``(tree, leaf_off, internal_off)``:
    a leaf ``o`` of segment ``s`` becomes ``leaf_off[s] + o``, an internal node ``o`` (``o >= n_leaf_s``) becomes
    ``internal_off[s] + o - n_leaf_s``.  ``locate`` maps back."""
    n_leaf_total = sum(t.n_leaf for t in trees)
    leaf_off, internal_off = [], []
    lo, io = 0, n_leaf_total
    for t in trees:
        leaf_off.append(lo)
        internal_off.append(io)
        lo += t.n_leaf
        io += t.n_nodes - t.n_leaf
    n = io
    if n >= 2**31:
        raise ValueError("node count exceeds int32 ordinals")
    parent_of = np.full(n, -1, dtype=np.int32)
    child_count = np.zeros(n, dtype=np.int32)
    prev_id = np.full(n, -1, dtype=np.int32)
    next_id = np.full(n, -1, dtype=np.int32)
    node_ids: Optional[List[str]] = [""] * n if all(t.node_ids is not None for t in trees) else None
    for s, t in enumerate(trees):
        nl = t.n_leaf

        def remap(a, nl=nl, s=s):
            a = a.astype(np.int64)
            return np.where(a < 0, -1, np.where(a < nl, a + leaf_off[s], a - nl + internal_off[s])).astype(np.int32)

        dst_leaf = slice(leaf_off[s], leaf_off[s] + nl)
        dst_int = slice(internal_off[s], internal_off[s] + t.n_nodes - nl)
        for src, dst in ((t.parent_of, parent_of), (t.prev_id, prev_id), (t.next_id, next_id)):
            m = remap(src)
            dst[dst_leaf] = m[:nl]
            dst[dst_int] = m[nl:]
        child_count[dst_leaf] = t.child_count[:nl]
        child_count[dst_int] = t.child_count[nl:]
        if node_ids is not None:
            node_ids[dst_leaf] = t.node_ids[:nl]
            node_ids[dst_int] = t.node_ids[nl:]
    return NodeTree(parent_of, child_count, prev_id, next_id, n_leaf_total, [], node_ids), leaf_off, internal_off


def locate(ordinal: int, trees: List[NodeTree], leaf_off: List[int], internal_off: List[int]):
    """Combined ordinal -> ``(segment, ordinal within that segment's own tree)``."""
    n_leaf_total = leaf_off[-1] + trees[-1].n_leaf
    offs = leaf_off if ordinal < n_leaf_total else internal_off
    s = int(np.searchsorted(np.asarray(offs), ordinal, side="right")) - 1
    return (s, ordinal - leaf_off[s]) if ordinal < n_leaf_total else (s, ordinal - internal_off[s] + trees[s].n_leaf)
