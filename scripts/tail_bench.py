"""Where the tail of a batch-1 query goes: every kernel after (and before) the scan timed on its own, 200 back-to-back
launches each on one stream (CUDA events), on a 10M-row index -- the shortlists are those of a real scan.  GPU box."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tensor_truth_b200 import _lib
from tensor_truth_b200.index import DeviceIndex, MergeResult
from tensor_truth_b200.synth import SynthCorpus

n = int(os.environ.get("ROWS", 10_000_000))
sc = SynthCorpus(n, 1024, 3, 1234, device="cuda")
corpus, inv = sc.rows(0, n)
q = sc.finish_queries(sc.queries(8, lookup=lambda t: corpus[t])).cuda()
idx = DeviceIndex(corpus, sc.tree, inv_norm=inv)
L, ptr = idx.lib, _lib.ptr
k, b = 10, 1
w = dict(idx._buffers(b, k))
mo = MergeResult(torch.empty((b, 2 * k), dtype=torch.int64, device="cuda"), torch.empty((b, 2 * k), dtype=torch.float64, device="cuda"),
                 torch.empty((b,), dtype=torch.int32, device="cuda"))
am = idx._am_args(0.5, mo)
r = idx.search(q[:1], k, out=w, am=am)   # fills the shortlists
torch.cuda.synchronize()
st = idx._stream()
n_cand = idx.n_lists * idx.kprime


def timed(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


res = {}
res["prepare_queries_us"] = timed(lambda: L.tt_prepare_queries(ptr(q), b, idx.dim, ptr(w["q_hi"]), ptr(w["q_lo"]), st))
res["stage2_prefilter_with_automerge_us"] = timed(lambda: idx._stage2(q[:1], b, w, n_cand, idx.n_lists, k, None, None, am, idx.eps))
res["stage2_prefilter_no_automerge_us"] = timed(lambda: idx._stage2(q[:1], b, w, n_cand, idx.n_lists, k, None, None, None, idx.eps))
os.environ["TT_NO_PREFILTER"] = "1"
res["stage2_full_rescore_with_automerge_us"] = timed(lambda: idx._stage2(q[:1], b, w, n_cand, idx.n_lists, k, None, None, am, idx.eps))
res["stage2_full_rescore_no_automerge_us"] = timed(lambda: idx._stage2(q[:1], b, w, n_cand, idx.n_lists, k, None, None, None, idx.eps))
del os.environ["TT_NO_PREFILTER"]
res["automerge_alone_us"] = timed(lambda: idx.automerge(r.ids, r.scores, out=mo))
x = torch.zeros(1, device="cuda")
res["empty_torch_kernel_us (launch floor)"] = timed(lambda: x.add_(1.0))
print(json.dumps(res))
