import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tensor_truth_b200.index import DeviceIndex
from tensor_truth_b200.synth import SynthCorpus
n = int(os.environ.get("ROWS", 10_000_000))
sc = SynthCorpus(n, 1024, 3, 1234, device="cuda")
corpus, inv = sc.rows(0, n)
q = sc.finish_queries(sc.queries(8, lookup=lambda t: corpus[t])).cuda()
idx = DeviceIndex(corpus, None, inv_norm=inv)
for b in (1, 2, 4, 8):
    r = idx.search_exact(q[:b], 10); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): idx.search_exact(q[:b], 10)
    e1.record(); torch.cuda.synchronize()
    print(f"exact scan B={b}: {e0.elapsed_time(e1)/5:.3f} ms")
