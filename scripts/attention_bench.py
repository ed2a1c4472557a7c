"""Attention of the rerank stage alone: tt_attention_varlen_bf16 (this package, tcgen05) against the library kernels
on the same packed QKV -- PyTorch's varlen_attn and, if importable, flash_attn's varlen kernel.  Shape: PAIRS pairs of
about TOKENS tokens, 16 heads x 64 (XLM-RoBERTa-large).  CUDA events, 100 launches each.  Prints one JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from tensor_truth_b200 import _lib

pairs, tokens = int(os.environ.get("PAIRS", 20)), int(os.environ.get("TOKENS", 384))
rng = np.random.default_rng(0)
lens = [int(tokens * rng.uniform(0.6, 1.0)) for _ in range(pairs)]
total, nh, h = sum(lens), 16, 1024
dev = torch.device("cuda:0")
qkv = torch.randn((total, 3 * h), device=dev).to(torch.bfloat16)
cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device=dev)
out = torch.empty((total, h), dtype=torch.bfloat16, device=dev)
L = _lib.lib()
s_max = max(lens)


def ours():
    _lib.check(L.tt_attention_varlen_bf16(qkv.data_ptr(), total, nh, 64, cu.data_ptr(), pairs, s_max, total // 128 + pairs, 0.125,
                                          out.data_ptr(), torch._C._cuda_getCurrentRawStream(0)))
    return out


arms = {"tt_attention_varlen_bf16": ours}
q, k, v = qkv.view(-1, 3, nh, 64).unbind(1)
try:
    from torch.nn.attention.varlen import varlen_attn

    arms["torch_varlen_attn"] = lambda: varlen_attn(q, k, v, cu, cu, s_max, s_max)
except Exception as exc:
    print("torch varlen_attn unavailable:", exc, file=sys.stderr)
try:
    from flash_attn import flash_attn_varlen_func

    qc, kc, vc = q.contiguous(), k.contiguous(), v.contiguous()
    arms["flash_attn_varlen"] = lambda: flash_attn_varlen_func(qc, kc, vc, cu, cu, s_max, s_max)
except Exception as exc:
    print("flash_attn unavailable:", exc, file=sys.stderr)

res, outs = {}, {}
for name, fn in arms.items():
    try:
        for _ in range(5):
            o = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            o = fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = {"us_per_call": e0.elapsed_time(e1) * 10.0}
        outs[name] = o.reshape(total, h).float()
    except Exception as exc:
        res[name] = {"error": f"{type(exc).__name__}: {exc}"}
for name in list(outs):
    if name != "tt_attention_varlen_bf16":
        res[name]["max_abs_diff_vs_ours"] = float((outs[name] - outs["tt_attention_varlen_bf16"]).abs().max())
flops = sum(4.0 * n * n * h for n in lens)
res["tt_attention_varlen_bf16"]["tflops"] = flops / (res["tt_attention_varlen_bf16"]["us_per_call"] * 1e-6) / 1e12
print(json.dumps({"workload": f"{pairs} packed sequences, {total} tokens, 16 heads x 64, bidirectional", **res}))
