"""Retrieval side of one chat turn, wired like RAGService.retrieve (services/rag_service.py:594-625): retriever.retrieve(str)
-> reranker.postprocess_nodes(nodes, query_bundle) -> cut to reranker_top_n.  Retriever = B200AutoMergingRetriever over a
10M x 1024 index, reranker = B200CrossEncoderRerank (XLM-RoBERTa-large shape, random weights), host strings in, node list
out.  The embedder is out of scope: a stand-in returns a prepared vector for the query string.  Prints one JSON line."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import transformers

from tensor_truth_b200.index import DeviceIndex
from tensor_truth_b200.rerank import B200CrossEncoder, B200CrossEncoderRerank, CrossEncoderWeights
from tensor_truth_b200.retriever import B200AutoMergingRetriever, B200VectorIndexRetriever, NodeTable
from tensor_truth_b200.schema import QueryBundle, TextNode
from tensor_truth_b200.synth import SynthCorpus

n = int(os.environ.get("ROWS", 10_000_000))
words = int(os.environ.get("WORDS_PER_LEAF", 96))   # ~128-token leaves; merged parents are 4-5x longer (cut at 512 tokens)
steps = int(os.environ.get("STEPS", 50))
sc = SynthCorpus(n, 1024, 3, 1234, device="cuda")
corpus, inv = sc.rows(0, n)
queries = sc.finish_queries(sc.queries(16, lookup=lambda t: corpus[t]))
lists = [row.tolist() for row in queries]
tree = sc.tree
n_leaf = tree.n_leaf


def text_of(ordinal):  # deterministic synthetic text; a parent's text is longer, like a 512 / 2048-token node
    k = words if ordinal < n_leaf else words * 4
    rng = np.random.default_rng(ordinal)
    return " ".join(f"w{int(x)}" for x in rng.integers(0, 30000, size=k))


class Embedder:
    vec = None

    def get_agg_embedding_from_queries(self, strs):
        return self.vec


emb = Embedder()
idx = DeviceIndex(corpus, tree, inv_norm=inv)
table = NodeTable(factory=lambda o: TextNode(id_=f"node-{o}", text=text_of(o), metadata={}))
retriever = B200AutoMergingRetriever(B200VectorIndexRetriever(idx, 10, emb, table), None)

torch.manual_seed(0)
cfg = transformers.XLMRobertaConfig(vocab_size=32000, hidden_size=1024, num_hidden_layers=24, num_attention_heads=16,
                                    intermediate_size=4096, max_position_embeddings=514, num_labels=1, type_vocab_size=1,
                                    pad_token_id=1, layer_norm_eps=1e-5)
model = transformers.XLMRobertaForSequenceClassification(cfg).eval()
enc = B200CrossEncoder(CrossEncoderWeights.from_hf_model(model, "cuda:0"), max_length=512)


def tokenize(pairs, max_length):  # stand-in for the SentencePiece tokenizer: one id per word
    def ids(s):
        return [3 + (int(w[1:]) % 31990) if w[1:].isdigit() else 3 + (hash(w) % 31990) for w in s.split()]

    return [([0] + ids(q) + [2, 2] + ids(d))[:max_length - 1] + [2] for q, d in pairs]


clock = {"tokenize": 0.0, "encoder": 0.0}


def timed_tokenize(pairs, max_length):
    t0 = time.perf_counter()
    out = tokenize(pairs, max_length)
    clock["tokenize"] += time.perf_counter() - t0
    return out


class TimedEncoder:
    max_length = enc.max_length

    def logits(self, toks):
        t0 = time.perf_counter()
        out = enc.logits(toks)
        out = out.cpu()                      # the scores are needed on the host: part of the encoder's latency
        clock["encoder"] += time.perf_counter() - t0
        return out


reranker = B200CrossEncoderRerank(TimedEncoder(), timed_tokenize, top_n=5)
reranker_top_n = 5


def turn(i):
    emb.vec = lists[i % len(lists)]
    question = f"what does document {i} say about tensor memory"
    t0 = time.perf_counter()
    nodes = retriever.retrieve(question)                                   # rag_service.py:594
    t1 = time.perf_counter()
    nodes = reranker.postprocess_nodes(nodes, query_bundle=QueryBundle(query_str=question))   # :611-616
    if reranker_top_n and len(nodes) > reranker_top_n:                     # :621-623
        nodes = nodes[:reranker_top_n]
    t2 = time.perf_counter()
    return nodes, t1 - t0, t2 - t1


for i in range(int(os.environ.get("WARM", 64))):  # covers the one-off CUDA-graph captures (retriever pipeline, reranker buckets)
    turn(i)
torch.cuda.synchronize()
tr, tk, cnt = 0.0, 0.0, 0
clock["tokenize"] = clock["encoder"] = 0.0
for i in range(steps):
    nodes, a, b = turn(1000 + i)
    tr, tk, cnt = tr + a, tk + b, cnt + len(nodes)
print(json.dumps({"workload": f"{n} x 1024 bf16 leaves, top-10 + auto-merge, then cross-encoder rerank to top-5 (XLM-RoBERTa-large "
                              f"shape, random weights, ~{words}-word leaves); host strings in, NodeWithScore list out",
                  "retrieve_ms": tr / steps * 1e3, "rerank_ms": tk / steps * 1e3,
                  "rerank_breakdown_ms": {"cross_encoder_forward": clock["encoder"] / steps * 1e3,
                                          "stand_in_tokenizer": clock["tokenize"] / steps * 1e3,
                                          "synthetic_text_and_python": (tk - clock["encoder"] - clock["tokenize"]) / steps * 1e3}, "total_ms": (tr + tk) / steps * 1e3,
                  "turns_per_s": steps / (tr + tk), "nodes_returned_avg": cnt / steps}))
