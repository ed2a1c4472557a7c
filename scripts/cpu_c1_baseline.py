"""The CPU baseline rows of BASELINE.md section 4: BASELINE configs[0] (C1: exact cosine top-10 over 100 000 x 1024
leaves, 64 queries, auto-merge over the 3-level tree) through the CPU restatement in its fast mode
(oracle/c/oracle_fast.c, fp32 AVX2 + OpenMP on every host thread; auto-merge in oracle/automerge.py) on the box's host
cores: the B = 1 loop and the B = 64 batch.  Prints one JSON line.  Test/bench infrastructure (it runs oracle/)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import oracle
from oracle import cport
from tensor_truth_b200.synth import make_small

cport.build()
threads = cport.use_all_host_threads()
tree, bits, inv, q = make_small(100_000, 64, dim=1024, levels=3, seed=1234)


def merge(ids, sc):
    return oracle.auto_merge([(int(o), float(s)) for o, s in zip(ids, sc) if o >= 0], tree.parent_of, tree.child_count,
                             tree.prev_id, tree.next_id)


def b1():
    for i in range(64):
        ids, sc = cport.fast_scan_topk(bits, inv, q[i:i + 1], 10)
        merge(ids[0], sc[0])


def b64():
    ids, sc = cport.fast_scan_topk(bits, inv, q, 10)
    for i in range(64):
        merge(ids[i], sc[i])


res = {}
for name, fn in (("b1_loop", b1), ("b64_batched", b64)):
    fn()
    ts = []
    for _ in range(7):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    res[name] = {"queries_per_s": 64 / float(np.median(ts)), "median_s_per_64_queries": float(np.median(ts))}
print(json.dumps({"config": "C1: 100000 x 1024 bf16 leaves, 64 queries, top-10 + auto-merge, 3-level tree", "threads": threads,
                  "host_cpus": os.cpu_count(), **res}))
