"""Per-query latency of retrieve_host (one synchronous caller) over the bench's 64-query pool: shows which queries leave
the typical path (many candidates inside stage 2's 2-eps window, repairs).  GPU box.  ROWS=10000000 by default."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tensor_truth_b200.index import DeviceIndex
from tensor_truth_b200.synth import SynthCorpus

n = int(os.environ.get("ROWS", 10_000_000))
sc = SynthCorpus(n, 1024, 3, 1234, device="cuda")
corpus, inv = sc.rows(0, n)
q = sc.finish_queries(sc.queries(64, lookup=lambda t: corpus[t])).cpu()
idx = DeviceIndex(corpus, sc.tree, inv_norm=inv)
for i in range(8):
    idx.retrieve_host(q[i:i + 1], 10)
lat = np.zeros((5, 64))
for rep in range(5):
    for i in range(64):
        t0 = time.perf_counter()
        idx.retrieve_host(q[i:i + 1], 10)
        lat[rep, i] = (time.perf_counter() - t0) * 1e6
best = lat.min(axis=0)
med = float(np.median(best))
print(f"median {med:.1f} us, min {best.min():.1f}, max {best.max():.1f}; queries more than 15 us above the median:")
for i in np.argsort(-best):
    if best[i] - med > 15:
        print(f"  query {i}: {best[i]:.1f} us (+{best[i] - med:.1f})")
print("retries", idx.retries, "deep", idx.deep_rescans, "fallbacks", idx.fallbacks)
