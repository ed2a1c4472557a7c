"""Times the stage-1 launch alone for several batch sizes / modes (tuning aid; run on the GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tensor_truth_b200 import _lib
from tensor_truth_b200.index import DeviceIndex
from tensor_truth_b200.synth import SynthCorpus

n = int(os.environ.get("ROWS", 10_000_000))
sc = SynthCorpus(n, 1024, 3, 1234, device="cuda")
corpus, inv = sc.rows(0, n)
q = sc.finish_queries(sc.queries(64, lookup=lambda t: corpus[t])).cuda()
idx = DeviceIndex(corpus, None, inv_norm=inv)
L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
cases = [(1, False), (8, False), (16, False), (32, False), (33, True), (48, True), (64, True), (16, True), (32, True)]
if len(sys.argv) > 1:
    cases = [(int(a.split(":")[0]), a.split(":")[1] == "hi") for a in sys.argv[1:]]
for b, hi_only in cases:
    w = idx._buffers(b, 10)
    qq = q[:b].contiguous()
    _lib.check(L.tt_prepare_queries(_lib.ptr(qq), b, 1024, _lib.ptr(w["q_hi"]), _lib.ptr(w["q_lo"]), st))

    def run():
        _lib.check(L.tt_scan_topk_bf16(_lib.ptr(idx.corpus), idx.n_rows, 1024, 1024, _lib.ptr(idx.inv_norm), _lib.ptr(w["q_hi"]),
                                       None if hi_only else _lib.ptr(w["q_lo"]), b, 32, 0, 0, _lib.ptr(w["cand_ids"]),
                                       _lib.ptr(w["cand_approx"]), _lib.ptr(w["cand_thresh"]), _lib.ptr(w["scan_ws"]),
                                       w["scan_ws"].numel(), st))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"B={b:3d} {'hi-only' if hi_only else 'hi+lo  '} {ms:8.3f} ms  {n * 2048 / ms / 1e6:7.0f} GB/s per call  {b / ms * 1e3:9.0f} q/s", flush=True)
