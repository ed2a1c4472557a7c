"""Several indexes as ONE segmented corpus on one GPU.

The reference opens one Chroma collection + docstore per module / session / project index and lets
``MultiIndexRetriever._retrieve_impl`` (/root/reference/src/tensortruth/rag_engine.py:406-461) fan every query
out over them on a thread pool -- m vector-store queries, m auto-merges, then ``_balance_top_k_per_index``
(:463-507).  Here the m leaf-embedding matrices are concatenated row-wise into one matrix in HBM and the m node
trees into one ordinal space (``tree.concat_trees``); one pass of the stage-1 kernel keeps a separate shortlist
per (segment, query) pair (``tt_scan_topk_bf16_segmented``), and stage 2 / auto-merge treat the n_seg * B
"virtual queries" v = s * B + b like any batch.  Per segment the result is what a ``DeviceIndex`` holding only
that segment returns (ids shifted by the segment's first row): exact top-k, ties by id, same certificate.
"""

from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import SCORE_COSINE, check, ptr
from .index import DeviceIndex, SearchResult, _as_device_corpus
from .tree import NodeTree, concat_trees, locate

class SegmentedIndex(DeviceIndex):
    """``DeviceIndex`` over the concatenation of several indexes; every search answers per segment."""

    COALESCE = False  # results come back as one list per (segment, query): concurrent callers are pipelined, not batched

    def __init__(self, corpora: Sequence, trees: Optional[Sequence[Optional[NodeTree]]] = None,
                 device: Optional[torch.device] = None, kprime: int = 32, score_mode: int = SCORE_COSINE):
        if not 1 <= len(corpora) <= _lib.MAX_SEGMENTS:
            raise ValueError(f"1..{_lib.MAX_SEGMENTS} segments, got {len(corpora)}")
        if not torch.cuda.is_available():
            raise RuntimeError("tensor_truth_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        parts = [_as_device_corpus(c, dev) for c in corpora]
        if len({p.dtype for p in parts}) != 1 or len({int(p.shape[1]) for p in parts}) != 1:
            raise ValueError("all segments must share dtype and embedding width")
        self.seg_trees = list(trees) if trees is not None and all(t is not None for t in trees) else None
        tree = None
        self.leaf_off = [0]
        for p in parts[:-1]:
            self.leaf_off.append(self.leaf_off[-1] + int(p.shape[0]))
        self.internal_off: List[int] = []
        if self.seg_trees is not None:
            for t, p in zip(self.seg_trees, parts):
                if t.n_leaf != int(p.shape[0]):
                    raise ValueError("a segment's tree must have one leaf per corpus row")
            tree, self.leaf_off, self.internal_off = concat_trees(self.seg_trees)
        super().__init__(torch.cat(parts, dim=0), tree, device=dev, kprime=kprime, score_mode=score_mode)
        self.n_seg = len(parts)
        self.seg_rows = [int(p.shape[0]) for p in parts]
        self._seg_end = (C.c_int64 * self.n_seg)(*np.cumsum(self.seg_rows).tolist())
        if self.dim % 128 or self.n_rows == 0:
            raise ValueError("the segmented scan needs dim % 128 == 0 and at least one row")
        if self.score_mode != SCORE_COSINE and not (self.norm_lo > 0.0 and self.norm_hi <= 1.05 * self.norm_lo):
            raise ValueError("chroma_l2_exp over a segmented corpus needs (near-)equal row norms")

    # ------------------------------------------------------------------ plumbing
    def _use_gemm(self, b: int, hi_only: bool = True) -> bool:
        return False

    def _result_rows(self, b: int) -> int:
        return self.n_seg * b

    def segment_of(self, ordinal: int):
        """Combined node ordinal -> (segment, ordinal within that segment's own tree)."""
        if self.seg_trees is None:
            s = int(np.searchsorted(np.asarray(self.leaf_off), ordinal, side="right")) - 1
            return s, ordinal - self.leaf_off[s]
        return locate(int(ordinal), self.seg_trees, self.leaf_off, self.internal_off)

    # ------------------------------------------------------------------ stage 1 + 2 over virtual queries
    def search(self, q: torch.Tensor, k: int, out: Optional[dict] = None, hi_only: Optional[bool] = None,
               xchg=None, am=None, row_filter=None) -> SearchResult:
        """Exact top-k of every segment for every query: rows v = s * B + b of the result belong to (segment s,
        query b); ids are rows of the concatenated corpus.  Asynchronous on the current stream."""
        if xchg is not None or hi_only:
            raise ValueError("SegmentedIndex.search: hi+lo queries on one GPU only")
        inv_norm = row_filter.inv_norm if row_filter is not None else self.inv_norm  # a gate over the concatenated rows
        q = self._check_queries(q)
        b = int(q.shape[0])
        vb = self.n_seg * b
        w = out if out is not None else self._buffers(vb, k, hi_only=False)
        q_rep = self._ws.get(("q_rep", vb))
        if q_rep is None:
            q_rep = self._ws[("q_rep", vb)] = torch.empty((vb, self.dim), dtype=torch.float32, device=self.device)
        cert = None
        if self.score_mode != SCORE_COSINE:
            cert = _lib.L2Cert(self.norm_lo, self.norm_hi, self.eps)
        L, st = self.lib, self._stream()
        n_cand = self.n_lists * self.kprime
        with self._on_device():
            q_rep.view(self.n_seg, b, self.dim).copy_(q.unsqueeze(0).expand(self.n_seg, b, self.dim))  # stage 2 reads q per virtual query
            check(L.tt_prepare_queries(ptr(q), b, self.dim, ptr(w["q_hi"]), ptr(w["q_lo"]), st))
            # one call; the library runs ceil(B / 8) passes, each writing its query columns of every segment
            check(L.tt_scan_topk_bf16_segmented(ptr(self.corpus), self.n_rows, self.dim, self._row_stride(self.corpus),
                                                ptr(inv_norm), ptr(w["q_hi"]), ptr(w["q_lo"]), b, self.kprime,
                                                self.id_base, self._seg_end, self.n_seg, ptr(w["cand_ids"]),
                                                ptr(w["cand_approx"]), ptr(w["cand_thresh"]), ptr(w["scan_ws"]),
                                                w["scan_ws"].numel(), st))
            self._stage2(q_rep, vb, w, n_cand, self.n_lists, k, None, cert, am, self.eps)
        return SearchResult(w["keys"], w["scores"], w["ids"], w["margin"], 0.0 if cert is not None else self.eps, False)

    def _repair(self, q, k, r: SearchResult, bad: torch.Tensor, hi_lo_first: bool, row_filter=None) -> None:
        """A (segment, query) pair whose certificate failed: the exact fp64 scan of that segment's rows."""
        b = int(q.shape[0])
        for v in bad.tolist():
            s, bq = divmod(int(v), b)
            lo = self.leaf_off[s]
            ex = self.search_exact(q[bq:bq + 1], k, rows=(lo, lo + self.seg_rows[s]), row_filter=row_filter)
            r.keys[v].copy_(ex.keys[0])
            r.scores[v].copy_(ex.scores[0])
            r.ids[v].copy_(ex.ids[0])
            self.fallbacks += 1

    def retrieve_host(self, q_host: torch.Tensor, k: int, ratio_thresh: float = 0.5, merge: bool = True):
        """Host queries [B, dim] in -> ``(ids [n_seg, B, w], scores [n_seg, B, w], lens [n_seg, B])`` (numpy), ids being
        ordinals of the combined tree (``segment_of`` maps them back).  Same pipeline, graph replay and certificate
        handling as ``DeviceIndex.retrieve_host``."""
        from .index import MAX_HOST_BATCH

        b = int(q_host.shape[0])
        if self.n_seg * b > MAX_HOST_BATCH:
            raise ValueError(f"at most {MAX_HOST_BATCH // self.n_seg} queries per call over {self.n_seg} segments")
        ids, scores, lens = super().retrieve_host(q_host, k, ratio_thresh, merge)
        return (ids.reshape(self.n_seg, b, -1), scores.reshape(self.n_seg, b, -1), lens.reshape(self.n_seg, b))

    def search_certified(self, q: torch.Tensor, k: int) -> SearchResult:
        q = self._check_queries(q)
        r = self.search(q, k)
        bad = torch.nonzero(~(r.margin > r.eps)).flatten()
        if bad.numel():
            self._repair(q, k, r, bad, hi_lo_first=False)
        return r
