"""Oracle, part 2: auto-merge over the hierarchical node tree.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py`` (parity unpinned).

Restates ``llama_index.core.retrievers.AutoMergingRetriever`` -- the object the
reference builds at /root/reference/src/tensortruth/rag_engine.py:641-643 (and
again at :676-679) with ``simple_ratio_thresh`` left at its default 0.5 -- over the
flat-array encoding of the docstore relations that
/root/reference/src/tensortruth/indexing/builder.py:385-442 writes:

* ``parent_of[o]``   ordinal of o's PARENT, -1 for level-0 nodes;
* ``child_count[o]`` ``len(node.child_nodes)`` (0 for leaves);
* ``prev_id[o]`` / ``next_id[o]``  PREVIOUS / NEXT sibling, -1 when absent.

The upstream method names are kept so a reader can put the two side by side:
``_fill_in_nodes``, ``_get_parents_and_merge``, ``_try_merging``, ``_retrieve``.
Scores are Python floats (float64) exactly as upstream carries them.

One interpreter-dependent detail is pinned here: upstream computes the parent score as
``sum(scores) / len(scores)``.  ``sum()`` over floats is a plain left-to-right float64
accumulation up to CPython 3.11 -- the version the reference's CI runs
(/root/reference/.github/workflows/tests.yml:16) -- and a compensated (Neumaier) sum from
3.12 on; the two can differ in the last ulp.  The oracle fixes the 3.11 meaning (explicit
left-to-right loop, ``_seq_sum``), so it does not depend on the interpreter running it.
"""

from __future__ import annotations

from typing import List, Sequence, Tuple

Pair = Tuple[int, float]


def _seq_sum(xs) -> float:
    acc = 0.0
    for x in xs:
        acc = acc + x
    return acc


def fill_in_nodes(nodes: List[Pair], prev_id: Sequence[int], next_id: Sequence[int]):
    """``_fill_in_nodes``: when exactly one sibling B sits between list-adjacent A and C
    (``A.next == C.prev``), insert B after A with the mean score.  No duplicate check."""
    out: List[Pair] = []
    changed = False
    n = len(nodes)
    for idx, (o, s) in enumerate(nodes):
        out.append((o, s))
        if idx >= n - 1:
            continue
        o2, s2 = nodes[idx + 1]
        nxt = int(next_id[o])
        if nxt != -1 and nxt == int(prev_id[o2]):
            changed = True
            out.append((nxt, (s + s2) / 2))
    return out, changed


def get_parents_and_merge(nodes: List[Pair], parent_of: Sequence[int], child_count: Sequence[int],
                          ratio_thresh: float = 0.5):
    """``_get_parents_and_merge``: group by parent (first-encounter order); where
    ``len(group) / len(parent.child_nodes) > ratio_thresh`` (strict) drop every copy of
    that parent's children and append the parent with the mean child score."""
    groups: dict = {}
    for o, s in nodes:
        p = int(parent_of[o])
        if p < 0:
            continue
        groups.setdefault(p, []).append(s)
    delete_parents = set()
    add: List[Pair] = []
    for p, scores in groups.items():
        num_children = int(child_count[p]) if int(child_count[p]) > 0 else 1
        ratio = len(scores) / num_children
        if ratio > ratio_thresh:
            delete_parents.add(p)
            add.append((p, _seq_sum(scores) / len(scores)))
    new_nodes = [(o, s) for (o, s) in nodes if int(parent_of[o]) not in delete_parents or int(parent_of[o]) < 0]
    new_nodes.extend(add)
    return new_nodes, len(delete_parents) > 0


def try_merging(nodes, parent_of, child_count, prev_id, next_id, ratio_thresh=0.5):
    nodes, ch0 = fill_in_nodes(nodes, prev_id, next_id)
    nodes, ch1 = get_parents_and_merge(nodes, parent_of, child_count, ratio_thresh)
    return nodes, (ch0 or ch1)


def auto_merge(nodes: List[Pair], parent_of, child_count, prev_id, next_id,
               ratio_thresh: float = 0.5, max_rounds: int = 64) -> List[Pair]:
    """``_retrieve`` after the base retriever returned ``nodes``: merge to a fixpoint, then a
    stable sort by score descending."""
    cur = [(int(o), float(s)) for o, s in nodes]
    cur, changed = try_merging(cur, parent_of, child_count, prev_id, next_id, ratio_thresh)
    rounds = 1
    while changed and rounds < max_rounds:
        cur, changed = try_merging(cur, parent_of, child_count, prev_id, next_id, ratio_thresh)
        rounds += 1
    cur.sort(key=lambda x: x[1], reverse=True)
    return cur
