"""Phase-growth sweep of the wide-batch stage 1 (TT_GEMM_GROWTH is read at every call).  ROWS=10000000 python scripts/gemm_tune.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tensor_truth_b200.index import DeviceIndex
from tensor_truth_b200.synth import SynthCorpus

n = int(os.environ.get("ROWS", 10_000_000))
sc = SynthCorpus(n, 1024, 3, 1234, device="cuda")
corpus, inv = sc.rows(0, n)
q = sc.finish_queries(sc.queries(4096, lookup=lambda t: corpus[t])).cuda()
idx = DeviceIndex(corpus, None, inv_norm=inv)
for k in (10, 100):
    for growth in (2, 3, 4, 6, 8):
        os.environ["TT_GEMM_GROWTH"] = str(growth)
        r = idx.search(q, k)
        torch.cuda.synchronize()
        ok = float((r.margin > r.eps).float().mean())
        idx.scan_events = []
        for _ in range(3):
            idx.search(q, k)
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(z) for a, z in idx.scan_events) / 3
        idx.scan_events = None
        print(f"k={k:4d} growth={growth}: stage1 {ms:8.2f} ms  {2.0 * 4096 * n * 1024 / ms / 1e9:7.1f} TFLOP/s  certified {ok:.4f}", flush=True)
