"""Auto-merge against the REAL upstream ``AutoMergingRetriever``.

``tests/golden/upstream_automerge.json`` is written by ``tests/golden/make_upstream_automerge_golden.py`` on a box that
has ``llama-index-core`` (this build image does not, and has no network).  When the file is present the CPU oracle --
and through it every CUDA parity test -- is pinned to upstream: same ids, float64 scores bit-equal, on the hand-built
cases and on trees parsed by the real ``HierarchicalNodeParser``.  Until someone commits it these rows stay
"parity unpinned" and this test says so instead of passing silently."""

import json
import os

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
PIN = os.path.join(HERE, "golden", "upstream_automerge.json")


def _load():
    if not os.path.exists(PIN):
        pytest.skip("parity unpinned: tests/golden/upstream_automerge.json is absent -- generate it with "
                    "tests/golden/make_upstream_automerge_golden.py on a box that has llama-index-core")
    with open(PIN) as f:
        return json.load(f)


def test_oracle_matches_upstream_on_handbuilt_cases():
    pin = _load()
    with open(os.path.join(HERE, "golden", "automerge_handbuilt.json")) as f:
        hand = {c["label"]: c for c in json.load(f)}
    assert pin["handbuilt"], "empty pin file"
    for got in pin["handbuilt"]:
        c = hand[got["label"]]
        arrs = [np.asarray(c[n], dtype=np.int32) for n in ("parent_of", "child_count", "prev_id", "next_id")]
        mine = oracle.auto_merge([(int(o), float(s)) for o, s in c["input"]], *arrs, c["ratio_thresh"])
        assert [[o, s] for o, s in mine] == got["got"], (got["label"], pin.get("llama_index_core_version"))


def test_oracle_matches_upstream_on_parsed_trees():
    pin = _load()
    if not pin.get("parsed"):
        pytest.skip(f"the pin file has no parsed-tree cases ({pin.get('parsed_error')})")
    t = pin["parsed"]["tree"]
    arrs = [np.asarray(t[n], dtype=np.int32) for n in ("parent_of", "child_count", "prev_id", "next_id")]
    for i, c in enumerate(pin["parsed"]["cases"]):
        mine = oracle.auto_merge([(int(o), float(s)) for o, s in c["input"]], *arrs, 0.5)
        assert [[o, s] for o, s in mine] == c["got"], i


def test_generator_script_is_self_consistent_without_upstream():
    """The script itself is exercised here as far as it can be without llama_index: it must import, find the hand-built
    cases, and fail with a clear message (not a traceback) when upstream is absent."""
    import subprocess
    import sys

    script = os.path.join(HERE, "golden", "make_upstream_automerge_golden.py")
    try:
        import llama_index.core  # noqa: F401
        pytest.skip("llama-index-core is installed here: run the script and commit its output instead")
    except ImportError:
        pass
    r = subprocess.run([sys.executable, script], capture_output=True, text=True)
    assert r.returncode != 0 and "llama-index-core is not importable" in (r.stderr + r.stdout)
    assert "Traceback" not in r.stderr
