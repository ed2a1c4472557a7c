"""Oracle, part 1: exact dense scan + deterministic top-k.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py`` (parity unpinned).

What the reference does at this step
(/root/reference/src/tensortruth/rag_engine.py:628-639): it opens the Chroma
collection ``"data"`` and asks ``index.as_retriever(similarity_top_k=k)`` for the k
nearest leaf embeddings of the query embedding.  ChromaDB answers from an
approximate HNSW index in squared-L2 space and ``ChromaVectorStore`` reports
``exp(-distance)``.  The *target* of that approximate search -- and what
BASELINE.json asks for -- is the exact brute-force result restated here.

Definitions pinned by this file (the CUDA path must reproduce them bit-for-bit
for ids, and to 1e-5 relative for scores):

* canonical values: the stored corpus values (bf16 or fp32), widened exactly to
  float64; the query as given (fp32), widened exactly to float64;
* ``dot = sum_i q_i * c_i``, ``qq = sum_i q_i^2``, ``nn = sum_i c_i^2`` in float64;
* ``SCORE_COSINE``:  ``key = score = float32(dot / (sqrt(qq) * sqrt(nn)))``
  (0 when either norm is 0);
* ``SCORE_CHROMA_L2_EXP``: ``d = float32(qq + nn - 2*dot)``, ``key = -d``,
  ``score = float32(exp(-float64(d)))``;
* order: ``key`` descending, ties -> smaller row ordinal first; fewer than k rows
  -> all rows (padded with id -1 / score -inf in the array form).
"""

from __future__ import annotations

import numpy as np

SCORE_COSINE = 0
SCORE_CHROMA_L2_EXP = 1

_BLOCK_ROWS = 16384


# --------------------------------------------------------------------------- bf16
def bf16_bits_to_f32(bits: np.ndarray) -> np.ndarray:
    """uint16 bf16 bit patterns -> float32 (exact)."""
    bits = np.ascontiguousarray(bits, dtype=np.uint16)
    return (bits.astype(np.uint32) << np.uint32(16)).view(np.float32)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> bf16 bit patterns, round-to-nearest-even (matches torch / cvt.rn.bf16.f32)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32)
    rounding = np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))
    out = ((u + rounding) >> np.uint32(16)).astype(np.uint16)
    nan = np.isnan(x)
    if nan.any():
        out = np.where(nan, np.uint16(0x7FC0), out)
    return out


def _rows_f64(corpus: np.ndarray, lo: int, hi: int) -> np.ndarray:
    blk = corpus[lo:hi]
    if blk.dtype == np.uint16:
        return bf16_bits_to_f32(blk).astype(np.float64)
    return np.asarray(blk, dtype=np.float64)


# --------------------------------------------------------------------------- scan
def exact_keys_f64(corpus: np.ndarray, queries: np.ndarray, score_mode: int = SCORE_COSINE):
    """All-rows exact scoring.  Returns ``(keys, scores)`` float32 ``[B, N]``.

    ``corpus``: ``[N, D]`` float32, or uint16 holding bf16 bits.  ``queries``: ``[B, D]`` float32.
    """
    q64 = np.atleast_2d(np.asarray(queries, dtype=np.float32)).astype(np.float64)
    n = corpus.shape[0]
    b = q64.shape[0]
    qq = np.einsum("bd,bd->b", q64, q64)
    keys = np.empty((b, n), dtype=np.float32)
    scores = np.empty((b, n), dtype=np.float32)
    for lo in range(0, n, _BLOCK_ROWS):
        hi = min(n, lo + _BLOCK_ROWS)
        c64 = _rows_f64(corpus, lo, hi)
        nn = np.einsum("nd,nd->n", c64, c64)
        dot = q64 @ c64.T  # [B, rows] float64
        if score_mode == SCORE_COSINE:
            den = np.sqrt(qq)[:, None] * np.sqrt(nn)[None, :]
            with np.errstate(divide="ignore", invalid="ignore"):
                s = np.where(den > 0.0, dot / den, 0.0)
            s32 = s.astype(np.float32)
            keys[:, lo:hi] = s32
            scores[:, lo:hi] = s32
        elif score_mode == SCORE_CHROMA_L2_EXP:
            d32 = (qq[:, None] + nn[None, :] - 2.0 * dot).astype(np.float32)
            keys[:, lo:hi] = -d32
            scores[:, lo:hi] = np.exp(-d32.astype(np.float64)).astype(np.float32)
        else:
            raise ValueError(f"unknown score_mode {score_mode}")
    return keys, scores


def _topk_by_key(keys: np.ndarray, k: int):
    """Indices of the k largest keys; order key desc, index asc.  NaN sorts last."""
    n = keys.shape[0]
    kk = np.where(np.isnan(keys), -np.inf, keys).astype(np.float32)
    if n <= k:
        cand = np.arange(n, dtype=np.int64)
    else:
        thr = np.partition(kk, n - k)[n - k]
        cand = np.nonzero(kk >= thr)[0].astype(np.int64)
    order = np.lexsort((cand, -kk[cand].astype(np.float64)))
    return cand[order][:k]


def exact_topk(corpus, queries, k: int, score_mode: int = SCORE_COSINE, id_base: int = 0):
    """Strict oracle: ``(ids int64 [B,k], scores float32 [B,k], keys float32 [B,k])``."""
    keys, scores = exact_keys_f64(corpus, queries, score_mode)
    b = keys.shape[0]
    ids = np.full((b, k), -1, dtype=np.int64)
    out_s = np.full((b, k), -np.inf, dtype=np.float32)
    out_k = np.full((b, k), -np.inf, dtype=np.float32)
    for i in range(b):
        sel = _topk_by_key(keys[i], k)
        ids[i, : sel.size] = sel + id_base
        out_s[i, : sel.size] = scores[i, sel]
        out_k[i, : sel.size] = keys[i, sel]
    return ids, out_s, out_k


def merge_topk_lists(keys_lists, ids_lists, k: int):
    """k-way merge of per-shard top-k lists for ONE query (row-sharded corpus).

    Same order rule as the scan: key desc, id asc.  Entries with id < 0 are padding.
    """
    keys = np.concatenate([np.asarray(x, dtype=np.float32).ravel() for x in keys_lists])
    ids = np.concatenate([np.asarray(x, dtype=np.int64).ravel() for x in ids_lists])
    ok = ids >= 0
    keys, ids = keys[ok], ids[ok]
    order = np.lexsort((ids, -keys.astype(np.float64)))[:k]
    return keys[order], ids[order]


def key_to_score(keys: np.ndarray, score_mode: int) -> np.ndarray:
    keys = np.asarray(keys, dtype=np.float32)
    if score_mode == SCORE_COSINE:
        return keys
    return np.exp(keys.astype(np.float64)).astype(np.float32)  # key = -d


# --------------------------------------------------------------------------- fast mode (timed CPU arm)
class FastCorpus:
    """fp32 copy of the corpus + inverse norms, prepared once (index-load time, untimed)."""

    def __init__(self, corpus: np.ndarray):
        import torch

        if corpus.dtype == np.uint16:
            c = torch.from_numpy(corpus.view(np.int16)).view(torch.bfloat16).to(torch.float32)
        else:
            c = torch.from_numpy(np.ascontiguousarray(corpus, dtype=np.float32))
        self.c = c
        self.nn = (c * c).sum(dim=1)
        self.inv_norm = torch.where(self.nn > 0, self.nn.rsqrt(), torch.zeros_like(self.nn))


def fast_topk(fc: FastCorpus, queries: np.ndarray, k: int, score_mode: int = SCORE_COSINE,
              block_rows: int = 65536):
    """CPU baseline in its fast mode: fp32 ``Q @ C^T`` on all host
    threads (MKL through torch), per-block ``topk``, then an id-ascending tie fix.

    This is the arm that gets *timed*; it scores in fp32, so it is not the parity oracle.
    """
    import torch

    q = torch.from_numpy(np.atleast_2d(np.asarray(queries, dtype=np.float32)))
    b = q.shape[0]
    qn = q.norm(dim=1)
    n = fc.c.shape[0]
    best_s = torch.full((b, 0), 0.0)
    best_i = torch.zeros((b, 0), dtype=torch.int64)
    for lo in range(0, n, block_rows):
        hi = min(n, lo + block_rows)
        dot = q @ fc.c[lo:hi].T
        if score_mode == SCORE_COSINE:
            s = dot * fc.inv_norm[lo:hi][None, :] / qn[:, None]
        else:
            s = -((qn * qn)[:, None] + fc.nn[lo:hi][None, :] - 2.0 * dot)
        kk = min(k, hi - lo)
        ts, ti = torch.topk(s, kk, dim=1)
        best_s = torch.cat([best_s, ts], dim=1)
        best_i = torch.cat([best_i, ti + lo], dim=1)
    ids = np.full((b, k), -1, dtype=np.int64)
    out = np.full((b, k), -np.inf, dtype=np.float32)
    bs, bi = best_s.numpy(), best_i.numpy()
    for i in range(b):
        order = np.lexsort((bi[i], -bs[i].astype(np.float64)))[:k]
        ids[i, : order.size] = bi[i, order]
        out[i, : order.size] = bs[i, order]
    if score_mode != SCORE_COSINE:
        out = np.exp(out.astype(np.float64)).astype(np.float32)
    return ids, out


# --------------------------------------------------------------------------- whole path
def retrieve(corpus, query, k, tree, score_mode: int = SCORE_COSINE, ratio_thresh: float = 0.5):
    """scan -> top-k -> auto-merge for one query.  Returns ``[(ordinal, score_float64), ...]``
    in the order ``AutoMergingRetriever._retrieve`` returns them."""
    from .automerge import auto_merge

    ids, scores, _ = exact_topk(corpus, np.asarray(query, dtype=np.float32)[None, :], k, score_mode)
    pairs = [(int(o), float(s)) for o, s in zip(ids[0], scores[0]) if o >= 0]
    return auto_merge(pairs, tree.parent_of, tree.child_count, tree.prev_id, tree.next_id, ratio_thresh)
