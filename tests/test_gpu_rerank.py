"""The reranker stage: B200CrossEncoder against the Hugging Face implementation of the same architecture
(XLMRobertaForSequenceClassification, one label) run in fp32 on the SAME randomly initialised weights.

Floating point, bf16 activations through up to 24 layers against an fp32 reference.  The yardstick is the library's own
bf16 forward pass of the same weights: the kernel path's worst logit error must not exceed max(0.03, 2 x the worst
error of Hugging Face's bf16 run), the reported score (sigmoid) must be within 0.01 + that error, and orderings must
agree wherever two reference scores differ by more than twice the tolerance.

Every test here needs a GPU:  python -m pytest tests -m gpu
"""

import numpy as np
import pytest
import torch

from tensor_truth_b200.rerank import B200CrossEncoder, B200CrossEncoderRerank, CrossEncoderWeights
from tensor_truth_b200.schema import NodeWithScore, QueryBundle, TextNode

pytestmark = pytest.mark.gpu
transformers = pytest.importorskip("transformers")


def _model(hidden, layers, heads, inter, vocab, seed):
    torch.manual_seed(seed)
    cfg = transformers.XLMRobertaConfig(vocab_size=vocab, hidden_size=hidden, num_hidden_layers=layers, num_attention_heads=heads,
                                        intermediate_size=inter, max_position_embeddings=514, num_labels=1, type_vocab_size=1,
                                        pad_token_id=1, layer_norm_eps=1e-5, hidden_dropout_prob=0.0,
                                        attention_probs_dropout_prob=0.0)
    m = transformers.XLMRobertaForSequenceClassification(cfg).eval()
    with torch.no_grad():  # a freshly initialised head gives logits ~0: spread them out so that ordering means something
        m.classifier.out_proj.weight.mul_(8.0)
    return m.cuda()


def _tokens(rng, n, vocab, lo, hi):
    out = []
    for _ in range(n):
        ln = int(rng.integers(lo, hi + 1))
        q = rng.integers(3, vocab, size=max(1, ln // 4)).tolist()
        d = rng.integers(3, vocab, size=max(1, ln - len(q) - 4)).tolist()
        out.append([0] + q + [2, 2] + d + [2])  # <s> query </s></s> passage </s>
    return out


def _hf_logits(model, toks, dtype=None):
    if dtype is not None:
        import copy

        model = copy.deepcopy(model).to(dtype)
    s = max(len(t) for t in toks)
    ids = torch.full((len(toks), s), 1, dtype=torch.long)
    for i, t in enumerate(toks):
        ids[i, :len(t)] = torch.tensor(t)
    ids = ids.cuda()
    with torch.no_grad():
        return model(input_ids=ids, attention_mask=(ids != 1).long()).logits.squeeze(-1).float()


@pytest.mark.parametrize("hidden,layers,heads,inter,n,lo,hi", [(256, 3, 4, 1024, 9, 6, 70), (1024, 6, 16, 4096, 16, 20, 300),
                                                                 (1024, 24, 16, 4096, 6, 100, 512)])
def test_cross_encoder_matches_hf_fp32(hidden, layers, heads, inter, n, lo, hi):
    vocab = 4000
    model = _model(hidden, layers, heads, inter, vocab, seed=layers)
    rng = np.random.default_rng(layers)
    toks = _tokens(rng, n, vocab, lo, hi)
    ref = _hf_logits(model, toks)
    enc = B200CrossEncoder(CrossEncoderWeights.from_hf_model(model, "cuda:0"))
    got = enc.logits(toks)
    torch.cuda.synchronize()
    hf_bf16_err = float((_hf_logits(model, toks, torch.bfloat16) - ref).abs().max())
    tol = max(0.03, 2.0 * hf_bf16_err)
    d = (got - ref).abs()
    assert float(d.max()) <= tol, (d.max().item(), hf_bf16_err, ref.tolist(), got.tolist())
    ps, pr = torch.sigmoid(got), torch.sigmoid(ref)
    assert float((ps - pr).abs().max()) < 0.01 + tol / 4  # d sigmoid / d logit <= 1/4
    order_ref = sorted(range(n), key=lambda i: -pr[i].item())
    for a, b in zip(order_ref[:-1], order_ref[1:]):
        if pr[a] - pr[b] > 2 * (0.01 + tol / 4):
            assert ps[a] > ps[b]
    print(f"layers={layers}: worst logit error {float(d.max()):.4f} (Hugging Face bf16: {hf_bf16_err:.4f})")
    # one pair alone gives the same logit as inside the batch (packing does not mix sequences)
    alone = enc.logits([toks[0]])
    assert abs(float(alone[0] - got[0])) < tol


def test_postprocess_nodes_semantics():
    vocab = 3000
    model = _model(256, 2, 4, 1024, vocab, seed=7)
    enc = B200CrossEncoder(CrossEncoderWeights.from_hf_model(model, "cuda:0"), max_length=64)

    def tokenize(pairs, max_length):  # a stand-in for the model's SentencePiece tokenizer: stable hash of the words
        def ids(s):
            return [3 + (hash_word(w) % (vocab - 3)) for w in s.split()]

        def hash_word(w):
            h = 0
            for ch in w:
                h = (h * 131 + ord(ch)) % 1_000_003
            return h

        return [([0] + ids(q) + [2, 2] + ids(d) + [2])[:max_length] for q, d in pairs]

    texts = [" ".join(f"w{(7 * i + j) % 50}" for j in range(5 + 3 * i)) for i in range(8)]
    nodes = [NodeWithScore(TextNode(id_=f"n{i}", text=t), 0.9 - 0.1 * i) for i, t in enumerate(texts)]
    qb = QueryBundle(query_str="w1 w2 w3 what is tensor memory")
    ref = torch.sigmoid(_hf_logits(model, tokenize([(qb.query_str, t) for t in texts], 64))).tolist()
    rr = B200CrossEncoderRerank(enc, tokenize, top_n=3, keep_retrieval_score=True)
    out = rr.postprocess_nodes(list(nodes), query_bundle=qb)
    assert len(out) == 3 and [n.score for n in out] == sorted((n.score for n in out), reverse=True)
    by_id = {f"n{i}": s for i, s in enumerate(ref)}
    assert all(abs(n.score - by_id[n.node.id_]) < 0.02 for n in out)
    best3 = sorted(by_id, key=lambda i: -by_id[i])[:3]
    if by_id[best3[2]] - sorted(by_id.values(), reverse=True)[3] > 0.02:
        assert {n.node.id_ for n in out} == set(best3)
    assert all("retrieval_score" in n.node.metadata for n in out)
    with pytest.raises(ValueError):
        rr.postprocess_nodes(list(nodes))          # upstream: "Missing query bundle in extra info."
    assert rr.postprocess_nodes([], query_bundle=qb) == []


def test_small_calls_replay_a_cuda_graph_and_match_the_eager_path(monkeypatch):
    """Interactive-size calls are launch-bound and are replayed as a bucketed CUDA graph (padding: dummy sequences);
    the replay must give what the eager path gives for the real pairs, call after call, for different token counts
    falling into the same bucket."""
    vocab = 4000
    model = _model(256, 3, 4, 1024, vocab, seed=11)
    rng = np.random.default_rng(11)
    enc = B200CrossEncoder(CrossEncoderWeights.from_hf_model(model, "cuda:0"))
    monkeypatch.setenv("TT_NO_GRAPH", "1")
    batches = [_tokens(rng, 6, vocab, 60, 100) for _ in range(5)]          # ~480 tokens each: one (768, 8) bucket
    eager = [enc.logits(b).clone() for b in batches]
    monkeypatch.delenv("TT_NO_GRAPH")
    got = [enc.logits(b).clone() for b in batches]
    torch.cuda.synchronize()
    assert any(g["graph"] is not None for g in enc._graphs.values()), "no bucket was captured"
    for a, b in zip(eager, got):
        assert float((a - b).abs().max()) < 2e-3
    ref = _hf_logits(model, batches[4])
    assert float((got[4] - ref).abs().max()) < 0.03
