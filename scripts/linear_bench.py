"""tt_linear_bf16 against torch.matmul (cuBLAS) on the reranker's layer shapes (tuning aid; GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tensor_truth_b200 import _lib

L = _lib.lib()
st = torch.cuda.current_stream().cuda_stream
for t, k, n in [(8192, 1024, 3072), (8192, 1024, 4096), (8192, 4096, 1024), (8192, 1024, 1024), (2048, 1024, 4096), (16384, 1024, 4096)]:
    x = torch.randn((t, k), device="cuda").to(torch.bfloat16)
    w = (torch.randn((n, k), device="cuda") / k ** 0.5).to(torch.bfloat16)
    b = torch.randn((n,), device="cuda")
    y = torch.empty((t, n), dtype=torch.bfloat16, device="cuda")

    def ours():
        _lib.check(L.tt_linear_bf16(_lib.ptr(x), t, k, _lib.ptr(w), n, _lib.ptr(b), None, 1, _lib.ptr(y), st))

    def cublas():
        return torch.nn.functional.gelu(torch.nn.functional.linear(x, w, b.to(torch.bfloat16)))

    res = {}
    for name, fn in (("tt_linear_bf16(+bias+gelu fused)", ours), ("torch linear + gelu", cublas)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        res[name] = (ms, 2.0 * t * k * n / ms / 1e9)
    print(f"T={t} K={k} N={n}: " + "; ".join(f"{nm}: {ms:.3f} ms {tf:.0f} TFLOP/s" for nm, (ms, tf) in res.items()), flush=True)
