"""Where the host time of one retrieve() goes (small corpus: the device work is ~50 us, the rest is host)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tensor_truth_b200.index import DeviceIndex
from tensor_truth_b200.retriever import B200AutoMergingRetriever, B200VectorIndexRetriever
from tensor_truth_b200.schema import QueryBundle
from tensor_truth_b200.synth import SynthCorpus

n = int(os.environ.get("ROWS", 100_000))
sc = SynthCorpus(n, 1024, 3, 1234, device="cuda")
corpus, inv = sc.rows(0, n)
q = sc.finish_queries(sc.queries(64, lookup=lambda t: corpus[t]))
idx = DeviceIndex(corpus, sc.tree, inv_norm=inv)
am = B200AutoMergingRetriever(B200VectorIndexRetriever(idx, 10), None)
lists = [row.tolist() for row in q]
call = lambda i: am.retrieve(QueryBundle(query_str="q", embedding=lists[i % 64]))
for i in range(20):
    call(i)
torch.cuda.synchronize()
N = 2000
t0 = time.perf_counter()
for i in range(N):
    call(i)
dt = (time.perf_counter() - t0) / N
print(f"{dt * 1e6:.1f} us per retrieve() at {n} rows")
if os.environ.get("PROFILE", "1") == "1":
    pr = cProfile.Profile()
    pr.enable()
    for i in range(N):
        call(i)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(18)
