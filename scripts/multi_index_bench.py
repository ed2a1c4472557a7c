"""Multi-index sessions: m indexes answered by (a) the reference caller's thread-pool fan-out over m
per-index B200 retrievers and (b) ONE pass over a SegmentedIndex.  Prints one JSON line.

    SEGMENTS=8 ROWS_PER_SEGMENT=250000 python scripts/multi_index_bench.py
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from oracle.multi_index import MultiIndexRetriever  # the caller restatement (pinned against the reference's class)
from tensor_truth_b200.index import DeviceIndex
from tensor_truth_b200.retriever import B200AutoMergingRetriever, B200MultiIndexRetriever, B200VectorIndexRetriever
from tensor_truth_b200.segmented import SegmentedIndex
from tensor_truth_b200.synth import SynthCorpus

m = int(os.environ.get("SEGMENTS", 8))
n = int(os.environ.get("ROWS_PER_SEGMENT", 250_000))
steps = int(os.environ.get("STEPS", 300))
parts, queries = [], []
for s in range(m):
    sc = SynthCorpus(n, 1024, 3, 100 + s, device="cuda")
    corpus, inv = sc.rows(0, n)
    parts.append((sc.tree, corpus))
    queries.append(sc.finish_queries(sc.queries(8, lookup=lambda t: corpus[t])))
q = torch.cat(queries)[torch.randperm(8 * m, generator=torch.Generator().manual_seed(0))]
lists = [row.tolist() for row in q]


class Fixed:
    vec = None

    def get_agg_embedding_from_queries(self, strs):
        return self.vec


emb = Fixed()
singles = [B200AutoMergingRetriever(B200VectorIndexRetriever(DeviceIndex(c, t), 10, emb), None) for t, c in parts]
fan = MultiIndexRetriever(singles, enable_cache=False)
seg = SegmentedIndex([c for _, c in parts], [t for t, _ in parts])
one = B200MultiIndexRetriever(seg, 10, emb, enable_cache=False)


def run(r, count):
    for i in range(count):
        emb.vec = lists[i % len(lists)]
        out = r.retrieve(f"q{i}")
    return out


same = True
for i in range(16):
    emb.vec = lists[i]
    a = [(x.node.metadata["_source_index"], x.node.id_, x.score) for x in fan.retrieve(f"p{i}")]
    b = [(x.node.metadata["_source_index"], x.node.id_, x.score) for x in one.retrieve(f"p{i}")]
    same &= a == b
res = {}
for name, r in (("fanout_threadpool_over_per_index_retrievers", fan), ("one_segmented_pass", one)):
    run(r, 20)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(r, steps)
    res[name] = {"ms_per_query": (time.perf_counter() - t0) / steps * 1e3}
bytes_total = m * n * 2048
res["one_segmented_pass"]["effective_GBps"] = bytes_total / (res["one_segmented_pass"]["ms_per_query"] / 1e3) / 1e9
print(json.dumps({"workload": f"{m} indexes x {n} x 1024 bf16 leaves, 3-level trees, top-10 + auto-merge per index, balance top_k_per_index, "
                              "host query in / NodeWithScore list out", "identical_results": same, **res,
                  "speedup": res["fanout_threadpool_over_per_index_retrievers"]["ms_per_query"] / res["one_segmented_pass"]["ms_per_query"]}))
