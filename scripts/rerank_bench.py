"""Rerank stage: B200CrossEncoder vs the Hugging Face model (bf16, SDPA) on the same weights.
bge-reranker-v2-m3's architecture (XLM-RoBERTa-large: 24 layers, hidden 1024, 16 heads, FFN 4096), random weights,
PAIRS (default 20 = k + merged parents at the default wiring) of TOKENS tokens each.  Prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import transformers

from tensor_truth_b200.rerank import B200CrossEncoder, CrossEncoderWeights

pairs, tokens = int(os.environ.get("PAIRS", 20)), int(os.environ.get("TOKENS", 384))
torch.manual_seed(0)
cfg = transformers.XLMRobertaConfig(vocab_size=32000, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
                                    intermediate_size=4096, max_position_embeddings=514, num_labels=1, type_vocab_size=1,
                                    pad_token_id=1, layer_norm_eps=1e-5, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
model = transformers.XLMRobertaForSequenceClassification(cfg).eval().cuda()
rng = np.random.default_rng(0)
toks = [[0] + rng.integers(3, 32000, size=int(tokens * rng.uniform(0.6, 1.0)) - 2).tolist() + [2] for _ in range(pairs)]
s = max(len(t) for t in toks)
ids = torch.full((pairs, s), 1, dtype=torch.long)
for i, t in enumerate(toks):
    ids[i, :len(t)] = torch.tensor(t)
ids = ids.cuda()
mask = (ids != 1).long()
enc = B200CrossEncoder(CrossEncoderWeights.from_hf_model(model, "cuda:0"))
with torch.no_grad():
    ref = model(input_ids=ids, attention_mask=mask).logits.squeeze(-1).float()
hf16 = model.to(torch.bfloat16)


def hf():
    with torch.no_grad():
        return hf16(input_ids=ids, attention_mask=mask).logits.squeeze(-1).float()


def ours():
    return enc.logits(toks)


res = {}
for name, fn in (("hf_bf16_sdpa", hf), ("b200_cross_encoder", ours)):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 10
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / n * 1e3
    res[name] = {"ms_per_call": ms, "pairs_per_s": pairs / ms * 1e3, "worst_logit_error_vs_fp32": float((out - ref).abs().max())}
total = sum(len(t) for t in toks)
flops = 24 * (2.0 * total * 1024 * (3 * 1024 + 1024 + 2 * 4096)) + 24 * sum(4.0 * len(t) * len(t) * 1024 for t in toks)
res["b200_cross_encoder"]["tflops"] = flops / (res["b200_cross_encoder"]["ms_per_call"] / 1e3) / 1e12
print(json.dumps({"workload": f"{pairs} (query, passage) pairs, {total} tokens packed ({s} padded), XLM-RoBERTa-large shape, random weights",
                  **res, "speedup": res["hf_bf16_sdpa"]["ms_per_call"] / res["b200_cross_encoder"]["ms_per_call"]}))
