"""Oracle vs the committed golden vectors, and the two restatements (numpy/Python vs C) vs each other.

Mirrors what a known-answer test for the path would look like in the reference, had it one
(it has none -- parity is unpinned against upstream)."""

import json
import os

import numpy as np
import pytest

import oracle
from oracle import cport
from tensor_truth_b200.synth import make_small
from tensor_truth_b200.tree import build_uniform_tree


def _cases(golden_dir):
    with open(os.path.join(golden_dir, "automerge_handbuilt.json")) as f:
        return json.load(f)


def test_handbuilt_automerge_python_and_c(golden_dir):
    cases = _cases(golden_dir)
    assert len(cases) >= 12
    for c in cases:
        pairs = [(int(o), float(s)) for o, s in c["input"]]
        tree = (c["parent_of"], c["child_count"], c["prev_id"], c["next_id"])
        exp = [(int(o), float(s)) for o, s in c["expected"]]
        assert oracle.auto_merge(pairs, *tree, ratio_thresh=c["ratio_thresh"]) == exp, c["label"]
        assert cport.auto_merge(pairs, *tree, ratio_thresh=c["ratio_thresh"]) == exp, c["label"]


def test_survey_worked_example(golden_dir):
    c = [x for x in _cases(golden_dir) if x["label"].startswith("survey_A4")][0]
    names = c["names"]
    got = [(names[o], s) for o, s in c["expected"]]
    assert [n for n, _ in got] == ["P", "e", "x"]
    assert got[0][1] == pytest.approx(0.7875, abs=1e-15)


def test_strict_ratio_threshold(golden_dir):
    by = {c["label"]: c for c in _cases(golden_dir)}
    c = by["ratio_exactly_half_no_merge"]
    assert [o for o, _ in c["expected"]] == [o for o, _ in c["input"]]  # 1/2 and 2/4 are not > 0.5
    c = by["ratio_just_over_half_merges"]
    assert c["names"][c["expected"][0][0]] == "R"


@pytest.mark.parametrize("tag,mode", [("cos", oracle.SCORE_COSINE), ("l2", oracle.SCORE_CHROMA_L2_EXP)])
@pytest.mark.parametrize("k", [10, 37])
def test_mini_scan_golden(golden_dir, tag, mode, k):
    g = np.load(os.path.join(golden_dir, "mini_scan.npz"))
    for fn in (oracle.exact_topk, cport.scan_topk):
        ids, sc, keys = fn(g["bits"], g["queries"], k, mode)
        assert (ids == g[f"{tag}_k{k}_ids"]).all()
        assert (sc == g[f"{tag}_k{k}_scores"]).all()
        assert (keys == g[f"{tag}_k{k}_keys"]).all()
    mi, ms = g[f"{tag}_k{k}_merged_ids"], g[f"{tag}_k{k}_merged_scores"]
    tree = (g["parent_of"], g["child_count"], g["prev_id"], g["next_id"])
    for i in range(ids.shape[0]):
        pairs = [(int(o), float(s)) for o, s in zip(ids[i], sc[i])]
        out = oracle.auto_merge(pairs, *tree)
        n = int((mi[i] >= 0).sum())
        assert [o for o, _ in out] == mi[i, :n].tolist()
        assert [s for _, s in out] == ms[i, :n].tolist()


def test_mini_scan_fp32_corpus_matches_bf16(golden_dir):
    g = np.load(os.path.join(golden_dir, "mini_scan.npz"))
    c32 = oracle.bf16_bits_to_f32(g["bits"])
    ids, sc, _ = oracle.exact_topk(c32, g["queries"], 10)
    assert (ids == g["cos_k10_ids"]).all() and (sc == g["cos_k10_scores"]).all()
    ids, sc, _ = cport.scan_topk(c32, g["queries"], 10)
    assert (ids == g["cos_k10_ids"]).all() and (sc == g["cos_k10_scores"]).all()


def test_bf16_roundtrip():
    rng = np.random.default_rng(0)
    x = rng.standard_normal(10000).astype(np.float32)
    b = oracle.f32_to_bf16_bits(x)
    import torch

    t = torch.from_numpy(x).to(torch.bfloat16)
    assert (t.view(torch.int16).numpy().view(np.uint16) == b).all()
    assert (oracle.bf16_bits_to_f32(b) == t.float().numpy()).all()


def test_ties_break_by_smaller_id_and_k_larger_than_n():
    rng = np.random.default_rng(1)
    base = rng.standard_normal((5, 32)).astype(np.float32)
    corpus = np.concatenate([base, base, base[:2]])  # rows 5..9 duplicate 0..4, 10..11 duplicate 0..1
    bits = oracle.f32_to_bf16_bits(corpus)
    q = rng.standard_normal((3, 32)).astype(np.float32)
    for fn in (oracle.exact_topk, cport.scan_topk):
        ids, sc, _ = fn(bits, q, 20)
        assert (ids[:, 12:] == -1).all() and np.isneginf(sc[:, 12:]).all()
        for b in range(3):
            row = ids[b, :12]
            assert sorted(row.tolist()) == list(range(12))
            for j in range(11):
                assert sc[b, j] > sc[b, j + 1] or (sc[b, j] == sc[b, j + 1] and row[j] < row[j + 1])
        # duplicates sit next to each other in id order
        pos0 = ids[0].tolist().index(0)
        assert ids[0, pos0:pos0 + 3].tolist() == [0, 5, 10]


def test_empty_corpus_and_zero_rows():
    bits = np.zeros((0, 16), np.uint16)
    q = np.ones((2, 16), np.float32)
    for fn in (oracle.exact_topk, cport.scan_topk):
        ids, sc, _ = fn(bits, q, 4)
        assert (ids == -1).all()
    z = np.zeros((3, 16), np.uint16)  # zero-norm rows score 0
    ids, sc, _ = oracle.exact_topk(z, q, 3)
    assert ids[0].tolist() == [0, 1, 2] and (sc == 0).all()


def test_merge_topk_lists_equals_global_topk(golden_dir):
    g = np.load(os.path.join(golden_dir, "mini_scan.npz"))
    bits, q = g["bits"], g["queries"]
    n = bits.shape[0]
    cuts = [0, 400, 401, 1000, n]
    for mode, tag in ((0, "cos"), (1, "l2")):
        parts = [oracle.exact_topk(bits[a:b], q, 10, mode, id_base=a) for a, b in zip(cuts[:-1], cuts[1:])]
        for i in range(q.shape[0]):
            keys, ids = oracle.merge_topk_lists([p[2][i] for p in parts], [p[0][i] for p in parts], 10)
            assert (ids == g[f"{tag}_k10_ids"][i]).all()
            assert (oracle.oracle.key_to_score(keys, mode) == g[f"{tag}_k10_scores"][i]).all()


def test_c1_config_golden(golden_dir):
    """BASELINE configs[0]: 100k x 1024, 64 queries, 3-level tree, top-10 + auto-merge."""
    g = np.load(os.path.join(golden_dir, "c1_expected.npz"))
    tree, bits, inv, q = make_small(100_000, 64, dim=1024, levels=3, seed=1234)
    import hashlib

    if hashlib.sha256(bits.tobytes()).hexdigest() != str(g["corpus_sha256"]):
        pytest.skip("torch CPU RNG stream differs from the one the fixture was generated with")
    assert (q == g["queries"]).all()
    ids, sc, _ = cport.scan_topk(bits, q, 10)
    assert (ids == g["ids"]).all() and (sc == g["scores"]).all()
    fast_ids, fast_sc = cport.fast_scan_topk(bits, inv, q, 10)
    assert (fast_ids == ids).mean() > 0.99 and np.abs(fast_sc - sc).max() < 1e-5
    n_merged = 0
    for i in range(64):
        pairs = [(int(o), float(s)) for o, s in zip(ids[i], sc[i])]
        out = oracle.auto_merge(pairs, tree.parent_of, tree.child_count, tree.prev_id, tree.next_id)
        n = int((g["merged_ids"][i] >= 0).sum())
        assert [o for o, _ in out] == g["merged_ids"][i, :n].tolist()
        assert [s for _, s in out] == g["merged_scores"][i, :n].tolist()
        n_merged += any(o >= tree.n_leaf for o, _ in out)
    assert n_merged > 32  # the synthetic tree is clustered so merges actually fire


def test_tree_invariants():
    t = build_uniform_tree(5000, levels=4, seed=3)
    t.validate()
    lo = t.level_offsets
    assert lo[0] == 0 and lo[1] == 5000 and len(lo) == 4
    # every non-top node has a parent one level up; child counts add up
    top = lo[-1]
    assert (t.parent_of[:top] >= 0).all() and (t.parent_of[top:] == -1).all()
    cc = np.bincount(t.parent_of[:top], minlength=t.n_nodes)
    assert (cc == t.child_count).all()
    assert ((t.child_count[lo[1]:] >= 1) & (t.child_count[lo[1]:] <= 6)).all()
    # prev/next only between siblings, and mutually consistent
    has_next = np.nonzero(t.next_id >= 0)[0]
    assert (t.prev_id[t.next_id[has_next]] == has_next).all()
    sib = has_next[has_next < top]
    assert (t.parent_of[sib] == t.parent_of[t.next_id[sib]]).all()
