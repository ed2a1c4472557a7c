"""Generates the golden vectors under tests/golden/ from the oracle.

The reference holds no golden vector, known-answer test or fixture for this path (its
``tests/fixtures/`` does not exist and every retrieval test mocks the retriever; SURVEY.md
section 8c), and the third-party packages that implement it are not installable here, so these
vectors come from ``oracle/`` (numpy/Python restatement), cross-checked against the independent C
restatement (``oracle/c``) at generation time.  Parity is therefore *unpinned* against upstream.

Run from the repo root:  python tests/golden/make_golden.py
"""

import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import cport  # noqa: E402
from tensor_truth_b200.synth import make_small  # noqa: E402
from tensor_truth_b200.tree import build_uniform_tree  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def hand_tree(spec):
    """spec: list of (name, parent_name or None, group) in ordinal order; prev/next link consecutive
    entries that share the same (parent, group)."""
    names = [s[0] for s in spec]
    idx = {n: i for i, n in enumerate(names)}
    n = len(spec)
    parent_of = [-1] * n
    child_count = [0] * n
    prev_id = [-1] * n
    next_id = [-1] * n
    for i, (name, par, _g) in enumerate(spec):
        if par is not None:
            parent_of[i] = idx[par]
            child_count[idx[par]] += 1
    for i in range(1, n):
        if spec[i][1] == spec[i - 1][1] and spec[i][2] == spec[i - 1][2] and (spec[i][1] is not None or spec[i][2] is not None):
            prev_id[i] = i - 1
            next_id[i - 1] = i
    return names, parent_of, child_count, prev_id, next_id


def handbuilt_cases():
    cases = []
    # SURVEY A.4 worked example: P{a,b,c,d}, Q{e,f}, R{x,y,z,w} under G (no parent).
    spec = [(c, "P", None) for c in "abcd"] + [(c, "Q", None) for c in "ef"] + [(c, "R", None) for c in "xyzw"] \
        + [("P", "G", None), ("Q", "G", None), ("R", "G", None), ("G", None, None)]
    names, *arrs = hand_tree(spec)
    ix = {n: i for i, n in enumerate(names)}

    def case(label, inp, ratio=0.5, tree=arrs, nm=names):
        idx = {n: i for i, n in enumerate(nm)}
        pairs = [(idx[a], s) for a, s in inp]
        out = oracle.auto_merge(pairs, *tree, ratio_thresh=ratio)
        out_c = cport.auto_merge(pairs, *tree, ratio_thresh=ratio)
        assert out == out_c, (label, out, out_c)
        cases.append(dict(label=label, names=nm, parent_of=tree[0], child_count=tree[1], prev_id=tree[2],
                          next_id=tree[3], ratio_thresh=ratio, input=[[a, s] for a, s in pairs],
                          expected=[[o, s] for o, s in out]))
        return out

    out = case("survey_A4_fill_in_duplicate_then_merge", [("a", .9), ("c", .8), ("e", .7), ("b", .6), ("x", .5)])
    assert [names[o] for o, _ in out] == ["P", "e", "x"] and abs(out[0][1] - 0.7875) < 1e-12
    case("ratio_exactly_half_no_merge", [("e", .9), ("x", .8), ("y", .7)])          # Q 1/2, R 2/4 -> neither > .5
    case("ratio_just_over_half_merges", [("x", .9), ("z", .8), ("w", .7), ("e", .1)])  # R 3/4
    case("two_level_cascade", [("a", .9), ("b", .8), ("c", .7), ("e", .65), ("f", .6), ("x", .2)])  # P,Q merge -> G 2/3
    case("fill_in_only_no_merge_needed", [("x", .9), ("z", .8)])                     # inserts y -> R 3/4 merges
    case("no_parent_nodes_untouched", [("G", .9), ("a", .5)])
    case("empty_input", [])
    case("single_node", [("a", .5)])
    case("tied_scores_stable_sort", [("e", .5), ("x", .5), ("a", .5)])
    case("duplicates_in_input", [("a", .9), ("a", .9), ("b", .1), ("e", .3)])        # P 3/4 counting the copy
    case("threshold_0.25", [("x", .9), ("y", .8), ("e", .7)], ratio=0.25)
    case("threshold_1.0_never", [("a", .9), ("b", .8), ("c", .7), ("d", .6)], ratio=1.0)
    # level-0 nodes of one document are prev/next linked but have no parent
    spec2 = [("r0", None, "doc"), ("r1", None, "doc"), ("r2", None, "doc"), ("s0", None, "doc2")]
    names2, *arrs2 = hand_tree(spec2)
    case("level0_fill_in_without_parent", [("r0", .9), ("r2", .7), ("s0", .1)], tree=arrs2, nm=names2)
    # parent with child_count 0 in the arrays (upstream: ``len(children) or 1``)
    names3 = ["u", "v", "W"]
    arrs3 = ([2, 2, -1], [0, 0, 0], [-1, 0, -1], [1, -1, -1])
    case("child_count_zero_counts_as_one", [("u", .4)], tree=arrs3, nm=names3)
    return cases


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def merged(tree, ids, scores):
    out = []
    for i in range(ids.shape[0]):
        pairs = [(int(o), float(s)) for o, s in zip(ids[i], scores[i]) if o >= 0]
        a = oracle.auto_merge(pairs, tree.parent_of, tree.child_count, tree.prev_id, tree.next_id)
        b = cport.auto_merge(pairs, tree.parent_of, tree.child_count, tree.prev_id, tree.next_id)
        assert a == b
        out.append(a)
    return out


def pack_merged(lists):
    n = max(len(x) for x in lists)
    ids = np.full((len(lists), n), -1, np.int64)
    sc = np.full((len(lists), n), np.nan, np.float64)
    for i, l in enumerate(lists):
        for j, (o, s) in enumerate(l):
            ids[i, j], sc[i, j] = o, s
    return ids, sc


def main():
    with open(os.path.join(HERE, "automerge_handbuilt.json"), "w") as f:
        json.dump(handbuilt_cases(), f, indent=1)

    # mini scan fixture: inputs AND outputs committed (N=1536 x D=64, 8 queries)
    tree, bits, inv, q = make_small(1536, 8, dim=64, levels=3, seed=77)
    out = {"bits": bits, "inv_norm": inv, "queries": q, "parent_of": tree.parent_of, "child_count": tree.child_count,
           "prev_id": tree.prev_id, "next_id": tree.next_id}
    for mode, tag in ((0, "cos"), (1, "l2")):
        for k in (10, 37):
            ids, sc, keys = oracle.exact_topk(bits, q, k, mode)
            ids_c, sc_c, keys_c = cport.scan_topk(bits, q, k, mode)
            assert (ids == ids_c).all() and (sc == sc_c).all() and (keys == keys_c).all()
            out[f"{tag}_k{k}_ids"], out[f"{tag}_k{k}_scores"], out[f"{tag}_k{k}_keys"] = ids, sc, keys
            mi, ms = pack_merged(merged(tree, ids, sc))
            out[f"{tag}_k{k}_merged_ids"], out[f"{tag}_k{k}_merged_scores"] = mi, ms
    np.savez_compressed(os.path.join(HERE, "mini_scan.npz"), **out)

    # C1 (BASELINE configs[0]): 100k x 1024, 64 queries, 3 levels, k=10 -- expected outputs + input hashes
    tree, bits, inv, q = make_small(100_000, 64, dim=1024, levels=3, seed=1234)
    ids, sc, keys = oracle.exact_topk(bits, q, 10, 0)
    ids_c, sc_c, _ = cport.scan_topk(bits, q, 10, 0)
    assert (ids == ids_c).all() and (sc == sc_c).all()
    mi, ms = pack_merged(merged(tree, ids, sc))
    t2 = build_uniform_tree(100_000, 3, 1234)
    assert (t2.parent_of == tree.parent_of).all()
    np.savez_compressed(os.path.join(HERE, "c1_expected.npz"), ids=ids, scores=sc, merged_ids=mi, merged_scores=ms,
                        queries=q, corpus_sha256=np.array(sha(bits)), tree_sha256=np.array(sha(tree.parent_of)),
                        corpus_head=bits[:64].copy())
    print("golden written:", os.listdir(HERE))


if __name__ == "__main__":
    main()
