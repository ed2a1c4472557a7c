"""The stage after the retrieval path: cross-encoder reranking.

The reference loads ``SentenceTransformerRerank(model="BAAI/bge-reranker-v2-m3", top_n=..., device=...)``
(/root/reference/src/tensortruth/services/model_manager.py:333-337) and runs
``postprocessor.postprocess_nodes(source_nodes, query_bundle=QueryBundle(query_str=...))`` on what the retriever
returned (services/rag_service.py:343-346): every (query, node text) pair goes through an XLM-RoBERTa
sequence-classification model, the node's score is overwritten with the model's sigmoid output, the list is sorted by it
and cut to ``top_n``.

``B200CrossEncoder`` is that model's forward pass on the B200 kernels of this package:

* all pairs of a call are PACKED (no padding) into one ``[T, hidden]`` activation matrix;
* input layer: ``tt_embed_layernorm_bf16`` (word + position + token-type embeddings, LayerNorm);
* every dense layer: ``tt_linear_bf16`` -- tcgen05 GEMM with bias / exact GELU / residual add fused in the epilogue
  (QKV as one [3H, H] GEMM, attention output + residual, FFN up + GELU, FFN down + residual);
* LayerNorms: ``tt_layernorm_bf16``;
* attention: ``tt_attention_varlen_bf16`` -- this package's packed variable-length kernel (one CTA per 128-row query
  tile and head; S = Q K^T and O += P V on tcgen05 with the accumulators in TMEM, softmax in registers; csrc/attention.cu);
* the classification head (dense + tanh + 1-unit projection on the <s> token): ``tt_cls_head_f32``.

No library kernel is left on the forward pass.  (``TT_RERANK_LIBRARY_ATTENTION=1`` routes attention through PyTorch's
``varlen_attn`` instead -- an A/B switch for scripts/rerank_bench.py, not a fallback: unsupported shapes raise.)

Weights come from a Hugging Face ``XLMRobertaForSequenceClassification`` state dict (``CrossEncoderWeights``); the
tokenizer is injected (``tokenize(pairs, max_length) -> List[List[int]]``: in a deployment the model's own
``AutoTokenizer``).  Neither the checkpoint nor the SentencePiece model is available in the build image, so tests and
benchmarks use randomly initialised weights of the same architecture and compare against the Hugging Face
implementation run in fp32 on the same weights.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, List, Optional, Sequence, Tuple

import os
import threading

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr

ACT_NONE, ACT_GELU = 0, 1
GRAPH_MAX_TOKENS = 4096    # calls with more packed tokens are GPU-bound: no graph
GRAPH_TOKEN_STEP = 256     # token-count granularity of the graph buckets
GRAPH_MAX_BUCKETS = 12     # captured graphs kept per encoder (each holds its activations: ~20 KB per token)

_LIBRARY_ATTENTION = bool(os.environ.get("TT_RERANK_LIBRARY_ATTENTION"))  # A/B switch for the bench script only
_varlen_attn = None
if _LIBRARY_ATTENTION:
    try:  # PyTorch >= 2.10: attention over packed sequences (cu_seqlens), no padding
        from torch.nn.attention.varlen import varlen_attn as _varlen_attn
    except Exception:  # pragma: no cover
        _varlen_attn = None


@dataclass
class _Layer:
    w_qkv: torch.Tensor
    b_qkv: torch.Tensor
    w_o: torch.Tensor
    b_o: torch.Tensor
    ln1_g: torch.Tensor
    ln1_b: torch.Tensor
    w_ff1: torch.Tensor
    b_ff1: torch.Tensor
    w_ff2: torch.Tensor
    b_ff2: torch.Tensor
    ln2_g: torch.Tensor
    ln2_b: torch.Tensor


class CrossEncoderWeights:
    """Device copies of an ``XLMRobertaForSequenceClassification`` (``num_labels == 1``) checkpoint in the layout the
    kernels want: GEMM operands bf16 ``[out, in]`` (the nn.Linear layout is already K-major), biases / LayerNorm
    parameters fp32, Q, K and V fused into one ``[3H, H]`` weight."""

    def __init__(self, state_dict, config, device):
        dev = torch.device(device)
        sd = {k: v for k, v in state_dict.items()}
        pre = "roberta." if any(k.startswith("roberta.") for k in sd) else ""

        def bf(name):
            return sd[name].detach().to(dev, torch.bfloat16).contiguous()

        def f32(name):
            return sd[name].detach().to(dev, torch.float32).contiguous()

        self.hidden = int(config.hidden_size)
        self.n_heads = int(config.num_attention_heads)
        self.inter = int(config.intermediate_size)
        self.eps = float(config.layer_norm_eps)
        self.pad_id = int(config.pad_token_id)
        if self.hidden % 256 or self.inter % 256 or self.hidden % self.n_heads:
            raise ValueError("hidden and intermediate sizes must be multiples of 256 (tt_linear_bf16 tile)")
        e = pre + "embeddings."
        self.word_emb, self.pos_emb, self.type_emb = bf(e + "word_embeddings.weight"), bf(e + "position_embeddings.weight"), \
            bf(e + "token_type_embeddings.weight")
        self.emb_ln_g, self.emb_ln_b = f32(e + "LayerNorm.weight"), f32(e + "LayerNorm.bias")
        self.layers: List[_Layer] = []
        for i in range(int(config.num_hidden_layers)):
            p = f"{pre}encoder.layer.{i}."
            a = p + "attention.self."
            w_qkv = torch.cat([sd[a + "query.weight"], sd[a + "key.weight"], sd[a + "value.weight"]], dim=0)
            b_qkv = torch.cat([sd[a + "query.bias"], sd[a + "key.bias"], sd[a + "value.bias"]], dim=0)
            self.layers.append(_Layer(
                w_qkv.detach().to(dev, torch.bfloat16).contiguous(), b_qkv.detach().to(dev, torch.float32).contiguous(),
                bf(p + "attention.output.dense.weight"), f32(p + "attention.output.dense.bias"),
                f32(p + "attention.output.LayerNorm.weight"), f32(p + "attention.output.LayerNorm.bias"),
                bf(p + "intermediate.dense.weight"), f32(p + "intermediate.dense.bias"),
                bf(p + "output.dense.weight"), f32(p + "output.dense.bias"),
                f32(p + "output.LayerNorm.weight"), f32(p + "output.LayerNorm.bias")))
        self.head_w1, self.head_b1 = f32("classifier.dense.weight"), f32("classifier.dense.bias")
        self.head_w2, self.head_b2 = f32("classifier.out_proj.weight"), f32("classifier.out_proj.bias")
        if self.head_w2.shape[0] != 1:
            raise ValueError("a reranker head has one output unit")
        self.device = dev

    @classmethod
    def from_hf_model(cls, model, device):
        return cls(model.state_dict(), model.config, device)


class B200CrossEncoder:
    """Packed forward pass of the cross-encoder; ``logits(token_lists)`` -> float32 ``[n_pairs]`` on the device."""

    def __init__(self, weights: CrossEncoderWeights, max_length: int = 512):
        if not torch.cuda.is_available():
            raise RuntimeError("tensor_truth_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.w = weights
        self.max_length = int(max_length)
        self.lib = _lib.lib()
        self._dev_index = weights.device.index if weights.device.index is not None else torch.cuda.current_device()
        self._pin: Optional[torch.Tensor] = None   # pinned staging buffer for the packed token ids / positions / offsets
        self._pin_free = None
        self._graphs: dict = {}
        self._lock = threading.Lock()              # staging buffers and graphs are per encoder: one call at a time

    # ---- kernels
    def _stream(self):
        return torch._C._cuda_getCurrentRawStream(self._dev_index)

    def _linear(self, x, w, b, residual=None, act=ACT_NONE):
        y = torch.empty((x.shape[0], w.shape[0]), dtype=torch.bfloat16, device=x.device)
        check(self.lib.tt_linear_bf16(ptr(x), int(x.shape[0]), int(x.shape[1]), ptr(w), int(w.shape[0]), ptr(b), ptr(residual),
                                      act, ptr(y), self._stream()))
        return y

    def _layernorm(self, x, g, b):
        y = torch.empty_like(x)
        check(self.lib.tt_layernorm_bf16(ptr(x), int(x.shape[0]), int(x.shape[1]), ptr(g), ptr(b), self.w.eps, ptr(y),
                                         self._stream()))
        return y

    # ---- attention: tt_attention_varlen_bf16 over the packed sequences (head_dim 64, pairs of up to 512 tokens)
    def _attention(self, qkv: torch.Tensor, cu: torch.Tensor, dest: torch.Tensor, n: int, s_max: int,
                   key_mask: torch.Tensor) -> torch.Tensor:
        h, nh = self.w.hidden, self.w.n_heads
        if not _LIBRARY_ATTENTION:
            total = int(qkv.shape[0])
            out = torch.empty((total, h), dtype=torch.bfloat16, device=qkv.device)
            check(self.lib.tt_attention_varlen_bf16(ptr(qkv), total, nh, h // nh, ptr(cu), n, int(s_max), total // 128 + n,
                                                    float((h // nh) ** -0.5), ptr(out), self._stream()))
            return out
        if _varlen_attn is not None:
            q, k, v = qkv.view(-1, 3, nh, h // nh).unbind(1)
            return _varlen_attn(q, k, v, cu, cu, s_max, s_max).reshape(-1, h)
        padded = torch.zeros((n * s_max, 3 * h), dtype=qkv.dtype, device=qkv.device)
        padded[dest] = qkv
        q, k, v = padded.view(n, s_max, 3, nh, h // nh).permute(2, 0, 3, 1, 4)
        out = torch.nn.functional.scaled_dot_product_attention(q, k, v, attn_mask=key_mask)
        return out.permute(0, 2, 1, 3).reshape(n * s_max, h)[dest].contiguous()

    def _forward(self, packed: torch.Tensor, total: int, n: int, s_max: int, lens_np=None) -> torch.Tensor:
        """Device forward over the packed int32 buffer ``[ids T | positions T | cu_seqlens n+1]`` -> logits ``[n]``."""
        w = self.w
        dev = w.device
        ids_d, pos_d, cu_d = packed[:total], packed[total:2 * total], packed[2 * total:2 * total + n + 1]
        dest_d = key_mask = None
        if _LIBRARY_ATTENTION and _varlen_attn is None:
            first_d = cu_d[:-1].long()
            lens_d = (cu_d[1:] - cu_d[:-1]).long()
            within = torch.arange(total, device=dev) - torch.repeat_interleave(first_d, lens_d)
            dest_d = within + torch.repeat_interleave(torch.arange(n, device=dev) * s_max, lens_d)
            key_mask = (torch.arange(s_max, device=dev)[None, :] < lens_d[:, None])[:, None, None, :]
        x = torch.empty((total, w.hidden), dtype=torch.bfloat16, device=dev)
        check(self.lib.tt_embed_layernorm_bf16(ptr(ids_d), ptr(pos_d), total, w.hidden, ptr(w.word_emb), ptr(w.pos_emb),
                                               ptr(w.type_emb), ptr(w.emb_ln_g), ptr(w.emb_ln_b), w.eps, ptr(x), self._stream()))
        for L in w.layers:
            qkv = self._linear(x, L.w_qkv, L.b_qkv)
            ctx = self._attention(qkv, cu_d, dest_d, n, s_max, key_mask)
            x = self._layernorm(self._linear(ctx, L.w_o, L.b_o, residual=x), L.ln1_g, L.ln1_b)
            hdn = self._linear(x, L.w_ff1, L.b_ff1, act=ACT_GELU)
            x = self._layernorm(self._linear(hdn, L.w_ff2, L.b_ff2, residual=x), L.ln2_g, L.ln2_b)
        logits = torch.empty((n,), dtype=torch.float32, device=dev)  # head on the <s> token of every pair
        check(self.lib.tt_cls_head_f32(ptr(x), ptr(cu_d), n, w.hidden, ptr(w.head_w1), ptr(w.head_b1), ptr(w.head_w2),
                                       ptr(w.head_b2), ptr(logits), self._stream()))
        return logits

    @staticmethod
    def _pack(hv: np.ndarray, token_lists, lens: np.ndarray, total: int, pad_id: int) -> None:
        """Fill ``hv`` = [ids total | positions total | cu_seqlens n+1] (int32) for the given (already clipped) lengths."""
        n = len(lens)
        first = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(lens, out=first[1:])
        for i, t in enumerate(token_lists):
            hv[first[i]:first[i + 1]] = t[:lens[i]] if len(t) > lens[i] else t
        # RoBERTa position ids start after the padding index: pad_id + 1 + offset inside the pair
        hv[total:2 * total] = np.arange(total) - np.repeat(first[:-1], lens) + (pad_id + 1)
        hv[2 * total:2 * total + n + 1] = first

    def _graph_bucket(self, total: int, n: int):
        """Small calls (the interactive case: ~10 nodes of a few hundred tokens) are launch-bound -- ~170 launches for
        ~1 ms of GPU work -- so their forward pass is replayed as ONE CUDA graph.  Graphs need fixed shapes: the call is
        padded to a bucket (tokens to a multiple of GRAPH_TOKEN_STEP, pairs to a power of two) with dummy one-token-or-
        longer sequences whose outputs are ignored; a bucket is captured on its second use, at most GRAPH_MAX_BUCKETS
        are kept.  Returns None when the call is too large (GPU-bound: eager is as fast) or graphs are disabled."""
        if (_LIBRARY_ATTENTION and _varlen_attn is None) or total > GRAPH_MAX_TOKENS or os.environ.get("TT_NO_GRAPH"):
            return None
        n_b = 2
        while n_b < n + 1:
            n_b *= 2
        t_b = (total + (n_b - n) + GRAPH_TOKEN_STEP - 1) // GRAPH_TOKEN_STEP * GRAPH_TOKEN_STEP
        if t_b - total > (n_b - n) * self.max_length:
            return None
        key = (t_b, n_b)
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= GRAPH_MAX_BUCKETS:
                self._graphs.pop(next(iter(self._graphs)))
            g = self._graphs[key] = {"uses": 0, "graph": None, "dead": False}
        g["uses"] += 1
        from .index import _CAPTURE_LOCK  # one capture at a time per process

        if g["graph"] is None and not g["dead"] and g["uses"] >= 2 and _CAPTURE_LOCK.acquire(blocking=False):
            try:
                dev = self.w.device
                host = torch.zeros(2 * t_b + n_b + 1, dtype=torch.int32).pin_memory()
                # a valid dummy problem for the capture run: n_b sequences of equal length
                lens0 = np.full(n_b, t_b // n_b, dtype=np.int64)
                lens0[: t_b - int(lens0.sum())] += 1
                self._pack(host.numpy(), [[0] * int(x) for x in lens0], lens0, t_b, self.w.pad_id)
                dev_in = torch.zeros_like(host, device=dev)
                side = torch.cuda.Stream(dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):  # warm-up outside the capture (library kernels may initialise lazily)
                    dev_in.copy_(host, non_blocking=True)
                    self._forward(dev_in, t_b, n_b, self.max_length)
                side.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side, capture_error_mode="thread_local"):
                    dev_in.copy_(host, non_blocking=True)
                    out = self._forward(dev_in, t_b, n_b, self.max_length)
                torch.cuda.current_stream(dev).wait_stream(side)
                # dev_in is what the graph's kernels read: it must live as long as the graph
                g.update(graph=graph, host=host, dev_in=dev_in, out=out, done=torch.cuda.Event())
            except Exception as exc:  # capture refused: this bucket stays eager
                import warnings

                warnings.warn(f"tensor_truth_b200: CUDA-graph capture of the cross-encoder failed ({exc}); staying eager")
                g["dead"] = True
                try:
                    torch.cuda.synchronize(self.w.device)
                except Exception:
                    pass
            finally:
                _CAPTURE_LOCK.release()
        return (g, t_b, n_b) if g["graph"] is not None else None

    def logits(self, token_lists: Sequence[Sequence[int]]) -> torch.Tensor:
        w = self.w
        n = len(token_lists)
        if n == 0:
            return torch.empty((0,), dtype=torch.float32, device=w.device)
        lens = np.fromiter((min(len(t), self.max_length) for t in token_lists), dtype=np.int64, count=n)
        if int(lens.min()) == 0:
            raise ValueError("empty token list")
        s_max, total = int(lens.max()), int(lens.sum())
        with self._lock, torch.cuda.device(w.device):
            bucket = self._graph_bucket(total, n)
            if bucket is not None:
                g, t_b, n_b = bucket
                if g.get("busy"):
                    g["done"].synchronize()          # the previous replay has consumed the staging buffer
                # dummy sequences soak up the padding: every one gets at least one token, none more than max_length
                pad = np.full(n_b - n, (t_b - total) // (n_b - n), dtype=np.int64)
                pad[: (t_b - total) - int(pad.sum())] += 1
                lens_b = np.concatenate([lens, pad])
                lists_b = list(token_lists) + [[w.pad_id + 2] * int(x) for x in pad]
                self._pack(g["host"].numpy(), lists_b, lens_b, t_b, w.pad_id)
                g["graph"].replay()
                g["done"].record()
                g["busy"] = True
                return g["out"][:n].clone()
            # eager: pack on the host in ONE int32 buffer [ids | positions | cu_seqlens] -> one pinned H2D copy
            need = 2 * total + n + 1
            if self._pin is None or self._pin.numel() < need:
                self._pin = torch.empty(max(need, 2 * self._pin.numel() if self._pin is not None else 0), dtype=torch.int32).pin_memory()
                self._pin_free = torch.cuda.Event()
            else:
                self._pin_free.synchronize()  # the previous call's H2D copy has consumed the staging buffer
            host = self._pin[:need]
            self._pack(host.numpy(), token_lists, lens, total, w.pad_id)
            packed = host.to(w.device, non_blocking=True)
            self._pin_free.record()
            return self._forward(packed, total, n, s_max)


class B200CrossEncoderRerank:
    """``SentenceTransformerRerank(model, top_n, device, keep_retrieval_score=False)``: ``postprocess_nodes(nodes,
    query_bundle)`` scores every (query, node text) pair, overwrites ``node.score`` with the sigmoid of the model's
    logit (what ``CrossEncoder.predict`` returns for a one-label model), sorts by it, descending, and keeps ``top_n``."""

    def __init__(self, encoder: B200CrossEncoder, tokenize: Callable[[List[Tuple[str, str]], int], List[List[int]]],
                 top_n: int = 2, keep_retrieval_score: bool = False):
        self.encoder, self.tokenize = encoder, tokenize
        self.top_n = int(top_n)
        self.keep_retrieval_score = bool(keep_retrieval_score)

    @staticmethod
    def _text(node: Any) -> str:
        inner = getattr(node, "node", node)
        try:
            from llama_index.core.schema import MetadataMode  # type: ignore

            return inner.get_content(metadata_mode=MetadataMode.EMBED)
        except Exception:
            return inner.get_content()

    def postprocess_nodes(self, nodes: List[Any], query_bundle: Optional[Any] = None, query_str: Optional[str] = None):
        if query_bundle is None and query_str is not None:
            from .schema import QueryBundle

            query_bundle = QueryBundle(query_str=query_str)
        if query_bundle is None:
            raise ValueError("Missing query bundle in extra info.")
        if len(nodes) == 0:
            return []
        pairs = [(query_bundle.query_str, self._text(n)) for n in nodes]
        logits = self.encoder.logits(self.tokenize(pairs, self.encoder.max_length))
        scores = torch.sigmoid(logits).cpu().tolist()
        for node, score in zip(nodes, scores):
            if self.keep_retrieval_score:
                node.node.metadata["retrieval_score"] = node.score
            node.score = float(score)
        return sorted(nodes, key=lambda x: -x.score if x.score else 0)[: self.top_n]
