"""Times the wide-batch (tensor-bound) stage 1 and the whole search for a few batch sizes (tuning aid; GPU box).

ROWS=10000000 python scripts/gemm_sweep.py 1024:10 4096:100 16384:100   # B:k
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tensor_truth_b200.index import DeviceIndex
from tensor_truth_b200.synth import SynthCorpus

n = int(os.environ.get("ROWS", 10_000_000))
sc = SynthCorpus(n, 1024, 3, 1234, device="cuda")
corpus, inv = sc.rows(0, n)
cases = [(int(a.split(":")[0]), int(a.split(":")[1])) for a in sys.argv[1:]] or [(1024, 10), (4096, 100)]
bmax = max(b for b, _ in cases)
q = sc.finish_queries(sc.queries(bmax, lookup=lambda t: corpus[t])).cuda()
idx = DeviceIndex(corpus, None, inv_norm=inv)
reps = int(os.environ.get("REPS", 3))
for b, k in cases:
    qq = q[:b].contiguous()
    r = idx.search(qq, k)
    torch.cuda.synchronize()
    ok = float((r.margin > r.eps).float().mean())
    min_margin = float(r.margin.min())
    idx.scan_events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        idx.search(qq, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    scan = sum(a.elapsed_time(z) for a, z in idx.scan_events) / reps
    idx.scan_events = None
    tf = 2.0 * b * n * 1024 / (scan * 1e-3) / 1e12
    print(f"B={b:6d} k={k:4d} gemm={idx._use_gemm(b)} search {ms:9.2f} ms  stage1 {scan:9.2f} ms  {tf:7.1f} TFLOP/s  "
          f"{b / ms * 1e3:9.0f} q/s  certified {ok:.4f}  min margin {min_margin:.5f} (eps {r.eps:.5f})", flush=True)
