"""Device-resident index: the leaf-embedding matrix + the node-tree relation arrays in HBM, and the
three-stage query pipeline over them (scan/shortlist -> exact re-score/top-k -> auto-merge).

This is what replaces, per index, the Chroma collection + docstore pair the reference opens at
/root/reference/src/tensortruth/rag_engine.py:628-639 and queries through
``index.as_retriever(similarity_top_k=k)`` / ``AutoMergingRetriever`` (:639-645).  All arithmetic runs
in ``libtt_b200.so`` (``include/tt_b200.h``); torch only owns the memory and the stream.

HBM layout (one shard):
  corpus     bf16 [n_rows, dim] row-major       streamed once per query batch by stage 1
  master     fp32 [n_rows, dim] (optional)      only when the stored embeddings are fp32; stage 2 reads it
  inv_norm   fp32 [n_rows]                      1/|c_r| of the bf16 rows (stage 1 epilogue)
  parent_of, child_count, prev_id, next_id   int32 [n_nodes]   (tree.py; replicated on every shard)
"""

from __future__ import annotations

import contextlib
import ctypes as C
import os
import threading
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import SCAN_AUTO, SCORE_COSINE, check, ptr
from .tree import NodeTree

# Certificate bounds on |approximate stage-1 score - exact cosine| ("Exactness"):
EPS_BF16_CORPUS = 2.5e-4   # corpus stored in bf16: hi/lo split residual + fp32 accumulation + inv_norm rounding
EPS_F32_CORPUS = 4.2e-3    # fp32 master scanned through a bf16 shadow: + 2^-8 corpus rounding
EPS_HI_ONLY = 3.95e-3      # added when the query travels as bf16 hi only (|q - bf16(q)| <= 2^-8 |q|)
HI_ONLY_ABOVE = 32         # batches larger than this scan hi-only first.  (17-32 queries hi-only through the GEMM-shaped scan were
                           # measured at 3.4 ms against 3.56 ms hi+lo: not worth giving up the tight certificate bound by default)
MAX_HOST_BATCH = 1024      # retrieve_host slices larger batches (shortlist workspace: 148 * K' * 12 B per query)
HILO_GEMM_FROM = 17        # hi+lo batches of 17-32 queries take the 64-column GEMM-shaped pass with (hi, lo) column pairs
GEMM_ABOVE = 33            # hi-only batches at least this large take the GEMM-shaped stage 1 (scan_gemm.cu): one corpus pass
                           # per 4096 queries, one list per query (64 queries: 3.04 ms at 10M rows vs 3.35 ms for the pair kernel)
# Queries per coalesced batch of concurrent retrieve_host callers: a scan pass costs the same up to 8.  Larger caps were
# tried with 16 / 32 caller threads (scripts/serving_probe.py): 16 gains nothing at 16 callers and 10 % at 32; with 32 --
# batches of 17-32 queries, i.e. the hi+lo GEMM pass on both lanes next to every other shape -- two of three runs at 10M
# rows ended in a device-side hang that fixed 32-query batches on two lanes do not reproduce (open issue).
COALESCE_MAX = 8
HOST_LANES = 2             # concurrent retrieve_host callers in flight per index (own stream, buffers, record and graph each)
GRAPH_AFTER = 3            # retrieve_host: eager calls of a (batch, k) shape before its pipeline is captured in a CUDA graph
GEMM_SLICE = 4096          # queries per GEMM-shaped corpus pass (candidate buffers: 16 K' * 8 B per query)


def gemm_kprime(k: int) -> int:
    """Shortlist length per query of the GEMM-shaped stage 1 (a single global list, so it is much deeper than k)."""
    forced = os.environ.get("TT_GEMM_KPRIME")  # tuning knob
    if forced and int(forced) >= k:
        return int(forced)
    return 128 if k <= 16 else 256 if k <= 128 else 512


_NULL_CTX = contextlib.nullcontext()
_CAPTURE_LOCK = threading.Lock()  # one CUDA-graph capture at a time per process


@contextlib.contextmanager
def capturing(graph, stream):
    """``torch.cuda.graph(graph, stream=stream)`` without its device-wide synchronisation on entry: whoever holds
    ``_CAPTURE_LOCK`` must never wait for the device.  With several host lanes in flight a device-wide wait includes
    another lane's merge kernel, which waits for a PEER's push, which that peer's thread may be unable to issue because it
    waits for the peer's own capture lock -- a lock-order cycle across ranks.  Nothing here needs the synchronisation: the
    shapes captured were run eagerly first, so the capture allocates nothing."""
    with torch.cuda.stream(stream):
        graph.capture_begin(capture_error_mode="thread_local")
        try:
            yield
        finally:
            graph.capture_end()


@dataclass
class SearchResult:
    keys: torch.Tensor      # float32 [B, k] ordering keys (cosine, or -squared-L2)
    scores: torch.Tensor    # float32 [B, k] reported scores (cosine, or exp(-squared-L2))
    ids: torch.Tensor       # int64   [B, k] global row ordinals, -1 padding
    margin: Optional[torch.Tensor]  # float32 [B] certificate margin (None for the exact scan)
    eps: float = 0.0        # the top-k of query b is proven exact iff margin[b] > eps
    hi_only: bool = False   # the scan saw bf16 hi halves of the queries only (the repair ladder starts with hi+lo)


@dataclass
class MergeResult:
    ids: torch.Tensor       # int64   [B, max_out] node ordinals, -1 padding
    scores: torch.Tensor    # float64 [B, max_out]
    lens: torch.Tensor      # int32   [B]


class StepGraph:
    """One query batch through the whole device pipeline, captured as a CUDA graph.  Inputs and outputs are
    device-resident and fixed: write the batch into ``q`` (stream-ordered, e.g. ``q.copy_(...)``), ``replay()``, read
    ``result`` (leaf top-k + certificate margins) / ``merged`` (auto-merged lists).  Produced by
    ``DeviceIndex.step_graph`` and ``ShardedIndex.step_graph``; replays go to the current stream."""

    def __init__(self, graph, q, result, merged, eps, extra=None):
        self.graph, self.q, self.result, self.merged, self.eps, self.extra = graph, q, result, merged, eps, extra

    def replay(self) -> None:
        self.graph.replay()


def capture_on_side_stream(device, fn):
    """``fn()`` once eagerly on the current stream (lazy initialisation; for a sharded index one real exchange, which
    every rank makes), then once more under capture on a side stream.  Returns ``(graph, fn's captured return value)``."""
    with _CAPTURE_LOCK:
        cur = torch.cuda.current_stream(device)
        fn()
        side = torch.cuda.Stream(device)
        side.wait_stream(cur)
        graph = torch.cuda.CUDAGraph()
        with capturing(graph, side):
            out = fn()
        cur.wait_stream(side)
    return graph, out


def _as_device_corpus(corpus, device):
    if isinstance(corpus, np.ndarray):
        if corpus.dtype == np.uint16:  # bf16 bit patterns
            t = torch.from_numpy(corpus.view(np.int16)).to(device).view(torch.bfloat16)
        else:
            t = torch.from_numpy(np.ascontiguousarray(corpus, dtype=np.float32)).to(device)
    else:
        t = corpus.to(device)
    if t.dtype not in (torch.bfloat16, torch.float32):
        raise TypeError(f"corpus dtype {t.dtype}: expected bfloat16 or float32")
    if t.dim() != 2:
        raise ValueError("corpus must be [n_rows, dim]")
    return t.contiguous()


class _HostRequest:
    """One ``retrieve_host`` call waiting for a lane (``DeviceIndex.retrieve_host``: coalescing of concurrent callers)."""

    __slots__ = ("q", "n", "key", "taken", "done", "result", "error")

    def __init__(self, q, n, key):
        self.q, self.n, self.key = q, n, key
        self.taken, self.done, self.result, self.error = False, threading.Event(), None, None


class _HostLane:
    """``DeviceIndex._host_lane()``: waits for a free host lane -- or, for a coalescable request, until some other caller
    has taken the request along in its batch, whichever comes first (then no lane is taken and ``-1`` is returned: a
    caller whose answer is being computed must not queue for a lane it does not need, or callers that re-enter quickly
    would starve it).  One condition variable guards the free-lane list and the pending requests."""

    __slots__ = ("idx", "req", "lane", "ctx")

    def __init__(self, idx, req=None):
        self.idx, self.req, self.lane, self.ctx = idx, req, -1, None

    def __enter__(self) -> int:
        idx, req = self.idx, self.req
        with idx._lane_cv:
            while True:
                if req is not None and req.taken:
                    return -1
                if idx._free_lanes:
                    self.lane = idx._free_lanes.pop(0)
                    break
                idx._lane_cv.wait()
        if self.lane > 0:
            st = idx._lane_streams.get(self.lane)
            if st is None:
                st = idx._lane_streams[self.lane] = torch.cuda.Stream(idx.device)
                st.wait_stream(torch.cuda.current_stream(idx.device))
            self.ctx = torch.cuda.stream(st)
            self.ctx.__enter__()
        return self.lane

    def __exit__(self, *exc) -> bool:
        if self.lane < 0:
            return False
        idx = self.idx
        try:
            if self.ctx is not None:
                self.ctx.__exit__(*exc)
        finally:
            with idx._lane_cv:
                idx._free_lanes.append(self.lane)
                idx._free_lanes.sort()       # lane 0 (the caller's own stream) first
                idx._lane_cv.notify_all()
        return False


class DeviceIndex:
    """One shard of one index on one GPU."""

    COALESCE = True  # concurrent retrieve_host callers may be served as one batch (SegmentedIndex: one result per segment -- off)

    def __init__(self, corpus, tree: Optional[NodeTree] = None, inv_norm: Optional[torch.Tensor] = None,
                 id_base: int = 0, device: Optional[torch.device] = None, kprime: int = 32,
                 variant: int = SCAN_AUTO, score_mode: int = SCORE_COSINE):
        if not torch.cuda.is_available():
            raise RuntimeError("tensor_truth_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.lib()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        stored = _as_device_corpus(corpus, self.device)
        if stored.dtype == torch.float32:
            self.master = stored
            self.corpus = stored.to(torch.bfloat16)
            self.eps = EPS_F32_CORPUS
        else:
            self.master = None
            self.corpus = stored
            self.eps = EPS_BF16_CORPUS
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.n_rows, self.dim = int(self.corpus.shape[0]), int(self.corpus.shape[1])
        if self.dim % 8:
            raise ValueError("dim must be a multiple of 8")
        self.id_base = int(id_base)
        self.kprime = int(kprime)
        self.variant = int(variant)
        self.score_mode = int(score_mode)
        with torch.cuda.device(self.device):
            self.n_lists = int(self.lib.tt_scan_num_lists(self.device.index or 0))
        if inv_norm is None:
            inv_norm = torch.empty(self.n_rows, dtype=torch.float32, device=self.device)
            step = 1 << 20
            for lo in range(0, self.n_rows, step):
                c = self.corpus[lo:lo + step].float()
                inv_norm[lo:lo + step] = (c * c).sum(dim=1).clamp_min(1e-30).rsqrt()
        self.inv_norm = inv_norm.to(self.device, torch.float32).contiguous()
        if self.master is not None and self.n_rows and not os.environ.get("TT_NO_STORE_EPS"):
            self.eps = min(EPS_F32_CORPUS, EPS_BF16_CORPUS + self._shadow_gap())
        # bounds on the stored rows' norms: what lets the cosine-ordered shortlist certify chroma_l2_exp mode
        self.norm_lo, self.norm_hi = 0.0, 0.0
        if self.n_rows:
            if self.master is not None:
                lo_v, hi_v = float("inf"), 0.0
                for a in range(0, self.n_rows, 1 << 20):
                    nrm = self.master[a:a + (1 << 20)].double().norm(dim=1)
                    lo_v, hi_v = min(lo_v, float(nrm.min())), max(hi_v, float(nrm.max()))
            else:
                nrm = 1.0 / self.inv_norm.double()
                lo_v, hi_v = float(nrm.min()), float(nrm.max())
            self.norm_lo, self.norm_hi = max(0.0, lo_v * (1 - 1e-5)), hi_v * (1 + 1e-5)
        self._ws: dict = {}
        # MultiIndexRetriever calls retrievers from a thread pool (rag_engine.py:420) and the web app serves requests
        # concurrently: ``retrieve_host`` callers are pipelined over HOST_LANES lanes (own stream, buffers, result record
        # and captured graph each), so one caller's tail + host work overlaps the next caller's corpus scan
        self._pending: list = []                                     # retrieve_host requests waiting for a lane
        self._lane_cv = threading.Condition()                        # guards _free_lanes, _pending and the requests' flags
        self._free_lanes = list(range(HOST_LANES))
        self._coalesce = not os.environ.get("TT_NO_COALESCE")
        self.coalesced = 0                                           # queries that rode along in another caller's batch
        self._lane_streams: dict = {}
        self._repair_lock = threading.Lock()  # the repair ladder's workspaces are shared; repairs are rare
        self.set_tree(tree)
        _lib.status_word(self._dev_index)  # device-side timeouts surface as TTError after the next host synchronisation
        self.fallbacks = 0             # queries whose certificate failed and were re-run through the exact scan
        self.retries = 0               # queries of a hi-only batch re-run through the hi+lo scan
        self.deep_rescans = 0          # queries re-run hi+lo with K' = 128 shortlists before the exact scan is tried
        self.scan_events = None        # bench hook: a list collects (start, end) CUDA events around every stage-1 launch

    def _shadow_gap(self) -> float:
        """fp32 store: what scanning the bf16 shadow instead of the master can cost any query, measured instead of
        budgeted.  Stage 1 scores row r as <q, c'_r * inv_norm_r> (c' the shadow row), the exact cosine is
        <q, c_r / |c_r|>; for a unit query the difference is at most g_r = |c'_r * inv_norm_r - c_r / |c_r||_2
        (Cauchy-Schwarz).  ``EPS_F32_CORPUS`` budgets the worst case 2^-8 for it; the maximum of g_r over the rows that are
        actually stored (one fp64 pass at load time, rounded up) is typically 2.5x smaller, which narrows both the
        certificate and stage 2's 2-eps pre-filter window."""
        g_max, step = 0.0, 1 << 17
        for a in range(0, self.n_rows, step):
            m = self.master[a:a + step].double()
            nrm = m.norm(dim=1, keepdim=True)
            unit = torch.where(nrm > 0, m / nrm.clamp_min(1e-300), torch.zeros_like(m))
            m = self.corpus[a:a + step].double() * self.inv_norm[a:a + step].double().unsqueeze(1)
            g = (m - unit).norm(dim=1)
            g = torch.where(nrm.squeeze(1) > 0, g, torch.zeros_like(g))  # an all-zero row scores 0 either way
            g_max = max(g_max, float(g.max()))
        return g_max * 1.001 + 2e-6

    # ------------------------------------------------------------------ tree
    def set_tree(self, tree: Optional[NodeTree]) -> None:
        """Install (or drop) the node tree.  Captured pipelines bake the old tree arrays' addresses in, so every cached
        graph / step graph is dropped with them."""
        held = self._hold_all_lanes() if hasattr(self, "_lane_cv") else None  # no retrieve_host call in flight meanwhile
        try:
            ws = getattr(self, "_ws", None)
            if ws:
                for key in [key for key in ws if isinstance(key, tuple) and key and key[0] in ("graph", "step")]:
                    del ws[key]
            self._set_tree_locked(tree)
        finally:
            if held is not None:
                self._release_lanes(held)

    def _hold_all_lanes(self):
        """Wait until no ``retrieve_host`` call is in flight and keep every lane (``set_tree``; tests)."""
        with self._lane_cv:
            while len(self._free_lanes) < HOST_LANES:
                self._lane_cv.wait()
            held, self._free_lanes = self._free_lanes, []
        return held

    def _release_lanes(self, held) -> None:
        with self._lane_cv:
            self._free_lanes = sorted(self._free_lanes + list(held))
            self._lane_cv.notify_all()

    def _host_lane(self, req=None):
        """Context: take a free host lane (blocks while all are busy).  Lane 0 runs on the caller's current stream, the
        others on a stream of their own, which is current inside the context.  ``with ... as lane``; with a coalescable
        request ``req`` the wait also ends -- with lane -1 -- when another caller has taken the request along."""
        return _HostLane(self, req)

    def _set_tree_locked(self, tree: Optional[NodeTree]) -> None:
        self.tree = tree
        if tree is None:
            self.parent_of = self.child_count = self.prev_id = self.next_id = None
            self.n_nodes = 0
            return
        tree.validate()
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(self.device)  # noqa: E731
        self.parent_of, self.child_count = up(tree.parent_of), up(tree.child_count)
        self.prev_id, self.next_id = up(tree.prev_id), up(tree.next_id)
        self.n_nodes = tree.n_nodes

    # ------------------------------------------------------------------ workspaces
    def _use_gemm(self, b: int, hi_only: bool = True) -> bool:
        """Wide hi-only batches: the tensor-bound regime, served by the GEMM-shaped stage 1."""
        return (hi_only and b >= int(os.environ.get("TT_GEMM_ABOVE", GEMM_ABOVE)) and self.variant in (SCAN_AUTO, _lib.SCAN_TCGEN05) and self.dim % 64 == 0
                and self.n_rows > 0 and not os.environ.get("TT_NO_GEMM"))

    def _use_gemm_hilo(self, b: int) -> bool:
        """17-32 hi+lo queries: the 64-column pass of the GEMM-shaped scan with every query on two MMA columns (hi, lo).
        The resident-query kernel needs 128 KB of shared memory for such a block and starves its TMA ring (3.56 ms per
        pass at 10M rows); the streaming pipeline keeps its bytes in flight."""
        return (int(os.environ.get("TT_HILO_GEMM_FROM", HILO_GEMM_FROM)) <= b <= 32 and self.variant in (SCAN_AUTO, _lib.SCAN_TCGEN05) and self.dim % 64 == 0
                and self.n_rows > 0 and not os.environ.get("TT_NO_GEMM") and not os.environ.get("TT_NO_GEMM_HILO"))

    def _buffers(self, b: int, k: int, slot: int = 0, hi_only: Optional[bool] = None, kprime: Optional[int] = None):
        """Workspace set for a (batch, k) shape; ``slot`` separates sets used concurrently on different streams;
        ``kprime`` overrides the per-CTA shortlist length (the repair ladder's deeper re-scan)."""
        hi = b > HI_ONLY_ABOVE if hi_only is None else hi_only
        hilo = (not hi) and kprime is None and self._use_gemm_hilo(b)
        gemm = (self._use_gemm(b, hi) and kprime is None) or hilo
        kprime = int(kprime or self.kprime)
        key = (b, k, slot, gemm, kprime, hilo)
        w = self._ws.get(key)
        if w is None and gemm:
            dev, kp = self.device, gemm_kprime(k)
            sl = min(b, int(os.environ.get("TT_GEMM_SLICE", GEMM_SLICE)))
            w = {
                "gemm": True, "kprime": kp, "slice": sl,
                "q_hi": torch.empty((b, self.dim), dtype=torch.bfloat16, device=dev),
                "q_lo": torch.empty((b, self.dim), dtype=torch.bfloat16, device=dev) if hilo else None,
                "cand_ids": torch.empty((b, kp), dtype=torch.int64, device=dev),
                "cand_approx": torch.empty((b, kp), dtype=torch.float32, device=dev),
                "cand_thresh": torch.empty((b, 1), dtype=torch.float32, device=dev),
                "ws": torch.zeros(max(8, int(self.lib.tt_rescore_fused_workspace_bytes(b, kp))), dtype=torch.uint8, device=dev),
                "gemm_ws": torch.empty(int(self.lib.tt_scan_gemm_workspace_bytes(sl, kp)), dtype=torch.uint8, device=dev),
                # more than one slice: odd slices run on a side stream with their own candidate buffers, so that one slice's
                # phase boundaries (drain, per-query cut, launch) are filled by the other slice's GEMM phases
                "gemm_ws2": (torch.empty(int(self.lib.tt_scan_gemm_workspace_bytes(sl, kp)), dtype=torch.uint8, device=dev)
                             if b > sl and not os.environ.get("TT_GEMM_ONE_STREAM") else None),
                "keys": torch.empty((b, k), dtype=torch.float32, device=dev),
                "scores": torch.empty((b, k), dtype=torch.float32, device=dev),
                "ids": torch.empty((b, k), dtype=torch.int64, device=dev),
                "margin": torch.empty((b,), dtype=torch.float32, device=dev),
                "rho": torch.empty((b,), dtype=torch.float32, device=dev),
            }
            self._ws[key] = w
        if w is None:
            dev, n_cand = self.device, self.n_lists * kprime
            w = {
                "kprime": kprime,
                "q_hi": torch.empty((b, self.dim), dtype=torch.bfloat16, device=dev),
                "q_lo": torch.empty((b, self.dim), dtype=torch.bfloat16, device=dev),
                "cand_ids": torch.empty((b, n_cand), dtype=torch.int64, device=dev),
                "cand_approx": torch.empty((b, n_cand), dtype=torch.float32, device=dev),
                "cand_thresh": torch.empty((b, self.n_lists), dtype=torch.float32, device=dev),
                "ws": torch.zeros(max(8, int(self.lib.tt_rescore_fused_workspace_bytes(b, n_cand))), dtype=torch.uint8, device=dev),
                "keys": torch.empty((b, k), dtype=torch.float32, device=dev),
                "scores": torch.empty((b, k), dtype=torch.float32, device=dev),
                "ids": torch.empty((b, k), dtype=torch.int64, device=dev),
                "margin": torch.empty((b,), dtype=torch.float32, device=dev),
                "rho": torch.empty((b,), dtype=torch.float32, device=dev),
                "scan_ws": torch.zeros(int(self.lib.tt_scan_workspace_bytes()), dtype=torch.uint8, device=dev),
            }
            self._ws[key] = w
        return w

    def _result_rows(self, b: int) -> int:
        """Result lists a batch of b queries produces (SegmentedIndex: one per segment and query)."""
        return b

    def _on_device(self):
        """Context that makes this index's GPU current -- a no-op object when it already is (the common case)."""
        if torch.cuda.current_device() == self._dev_index:
            return _NULL_CTX
        return torch.cuda.device(self.device)

    def _row_stride(self, t: torch.Tensor) -> int:
        return max(int(t.stride(0)), self.dim)  # an empty tensor reports stride 0

    def _stream(self):
        """Raw handle of torch's current stream on this index's GPU (the C call: ``torch.cuda.current_stream()`` costs
        ~15 us of Python per call, three of them per query)."""
        return torch._C._cuda_getCurrentRawStream(self._dev_index)

    def _check_queries(self, q: torch.Tensor) -> torch.Tensor:
        if q.dim() != 2 or q.shape[1] != self.dim:
            raise ValueError(f"queries must be [B, {self.dim}], got {tuple(q.shape)}")
        if q.device != self.device or q.dtype != torch.float32 or not q.is_contiguous():
            q = q.to(self.device, torch.float32).contiguous()
        return q

    # ------------------------------------------------------------------ stage 1 + 2
    def _am_args(self, ratio_thresh: float, out: "MergeResult", max_rounds: int = 64):
        """``tt_automerge_args_t`` for the kernels that run stage 3 as their tail (tree arrays + the outputs in ``out``)."""
        if self.tree is None:
            raise RuntimeError("this index has no node tree: auto-merge is not available")
        return _lib.AutomergeArgs(ptr(self.parent_of), ptr(self.child_count), ptr(self.prev_id), ptr(self.next_id),
                                  self.n_nodes, float(ratio_thresh), int(max_rounds), ptr(out.ids), ptr(out.scores),
                                  ptr(out.lens), int(out.ids.shape[1]))

    def _stage2(self, q, b, w, n_cand, n_lists, k, xchg=None, cert=None, am=None, eps=0.0):
        """Exact re-score of the shortlist + top-k selection in ONE launch (``tt_rescore_topk_fused``), which then pushes
        the record to the peer ranks (``xchg``) or runs the auto-merge on it (``am``).  In cosine mode only the
        candidates within ``2 eps`` of the k-th best approximate score are re-scored (``eps``: the stage-1 error bound
        the certificate is checked against) -- a candidate further below cannot reach the exact top-k."""
        src = self.master if self.master is not None else self.corpus
        window = 2.0 * float(eps) if (cert is None and not os.environ.get("TT_NO_PREFILTER")) else 0.0
        check(self.lib.tt_rescore_topk_fused(ptr(src), _lib.DTYPE_F32 if self.master is not None else _lib.DTYPE_BF16,
                                             self.n_rows, self.dim, self._row_stride(src), self.id_base, ptr(q), b,
                                             ptr(w["cand_ids"]), n_cand, ptr(w["cand_thresh"]), n_lists, k, self.score_mode,
                                             ptr(w["keys"]), ptr(w["scores"]), ptr(w["ids"]), ptr(w["margin"]),
                                             ptr(w["ws"]), w["ws"].numel(), C.byref(xchg) if xchg is not None else None,
                                             C.byref(cert) if cert is not None else None,
                                             C.byref(am) if am is not None else None, ptr(w["cand_approx"]), window,
                                             self._stream()))

    def row_filter(self, eligible, key=None):
        """A ``filters.RowFilter`` for this index (bool ``[n_rows]`` eligibility -> gated ``inv_norm`` in HBM), cached by
        ``key`` when one is given (e.g. a canonical form of the filter)."""
        from .filters import RowFilter

        if key is not None:
            hit = self._ws.get(("row_filter", key))
            if hit is not None:
                return hit
        rf = RowFilter(self, eligible, key)
        if key is not None:
            self._ws[("row_filter", key)] = rf
        return rf

    def search(self, q: torch.Tensor, k: int, out: Optional[dict] = None, hi_only: Optional[bool] = None,
               xchg=None, am=None, row_filter=None) -> SearchResult:
        """Shortlist scan + exact re-score.  Asynchronous on the current stream; ``margin[b] > result.eps``
        certifies that query b's top-k is the exact one (``search_certified`` acts on it).
        ``hi_only``: send the queries through the tensor cores as bf16 hi halves only (twice the queries per
        corpus pass, wider certificate); default: batches above ``HI_ONLY_ABOVE``.
        ``xchg`` (``_lib.Exchange``): row-sharded corpus -- the selecting kernel also pushes this shard's top-k
        record to every peer rank (sharded.py).
        ``am`` (``_am_args``): single shard -- the selecting kernel also auto-merges the list it selected.
        ``row_filter`` (``filters.RowFilter``): only the rows it admits exist for this search (metadata filters)."""
        q = self._check_queries(q)
        b = int(q.shape[0])
        if hi_only is None:
            hi_only = b > HI_ONLY_ABOVE
        w = out if out is not None else self._buffers(b, k, hi_only=hi_only)
        inv_norm = row_filter.inv_norm if row_filter is not None else self.inv_norm
        l2 = self.score_mode != SCORE_COSINE
        if l2 and not (self.norm_lo > 0.0 and self.norm_hi <= 1.05 * self.norm_lo):
            # chroma_l2_exp on rows of clearly unequal norm: cosine order says little about squared-L2 order, the
            # certificate below would refuse every query -- let the exact fp64 scan answer directly.
            ex = self.search_exact(q, k, out=w, row_filter=row_filter)
            w["margin"].fill_(float("inf"))
            if xchg is not None:
                raise ValueError("chroma_l2_exp over a row-sharded corpus needs (near-)equal row norms")
            if am is not None:
                check(self.lib.tt_automerge(ptr(ex.ids), ptr(ex.scores), b, k, am.parent_of, am.child_count, am.prev_id,
                                            am.next_id, am.n_nodes, am.ratio_thresh, am.max_rounds, am.out_ids,
                                            am.out_scores, am.out_len, am.max_out, self._stream()))
            return SearchResult(ex.keys, ex.scores, ex.ids, w["margin"], 0.0)
        eps = self.eps + (EPS_HI_ONLY if hi_only else 0.0)
        cert = None
        if l2:  # (near-)unit-norm rows: the cosine bound of a dropped row bounds its L2 key (tt_l2_cert_t)
            cert = _lib.L2Cert(self.norm_lo, self.norm_hi, eps)
        L, st = self.lib, self._stream()
        gemm = bool(w.get("gemm")) and (hi_only or w.get("q_lo") is not None)
        kprime = w.get("kprime", self.kprime)
        n_cand, n_lists = (kprime, 1) if gemm else (self.n_lists * kprime, self.n_lists)
        with self._on_device():
            credit = hi_only and not os.environ.get("TT_NO_CERT_CREDIT")
            if credit:  # hi-only: also measure what the scan will not see of each query (certificate credit, below)
                rho = w.get("rho")
                if rho is None:
                    rho = w["rho"] = torch.empty((b,), dtype=torch.float32, device=self.device)
                check(L.tt_prepare_queries_rho(ptr(q), b, self.dim, ptr(w["q_hi"]), ptr(w["q_lo"]), ptr(rho), st))
            else:
                check(L.tt_prepare_queries(ptr(q), b, self.dim, ptr(w["q_hi"]), ptr(w["q_lo"]), st))
            if self.scan_events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            if gemm:  # the tensor-bound regime: GEMM-shaped scan, one shortlist per query, GEMM_SLICE queries per corpus pass
                ws2 = w.get("gemm_ws2")
                side = None
                if ws2 is not None:  # fork: odd slices go to a side stream (joined below)
                    side = self._ws.get("side_stream")
                    if side is None:
                        side = self._ws["side_stream"] = torch.cuda.Stream(self.device)
                    cur = torch.cuda.current_stream(self.device)
                    fork = torch.cuda.Event()
                    fork.record(cur)
                    side.wait_event(fork)
                for i, a in enumerate(range(0, b, w["slice"])):
                    n = min(w["slice"], b - a)
                    odd = side is not None and (i & 1)
                    gws = ws2 if odd else w["gemm_ws"]
                    check(L.tt_scan_gemm_topk_bf16(ptr(self.corpus), self.n_rows, self.dim, self._row_stride(self.corpus),
                                                   ptr(inv_norm), ptr(w["q_hi"][a:]), None if hi_only else ptr(w["q_lo"][a:]), n,
                                                   n_cand, self.id_base,
                                                   ptr(w["cand_ids"][a:]), ptr(w["cand_approx"][a:]), ptr(w["cand_thresh"][a:]),
                                                   ptr(gws), gws.numel(), side.cuda_stream if odd else st))
                if side is not None:  # join
                    join = torch.cuda.Event()
                    join.record(side)
                    cur.wait_event(join)
            else:
                check(L.tt_scan_topk_bf16(ptr(self.corpus), self.n_rows, self.dim, self._row_stride(self.corpus), ptr(inv_norm),
                                          ptr(w["q_hi"]), None if hi_only else ptr(w["q_lo"]), b, kprime, self.id_base,
                                          self.variant,
                                          ptr(w["cand_ids"]), ptr(w["cand_approx"]), ptr(w["cand_thresh"]),
                                          ptr(w["scan_ws"]), w["scan_ws"].numel(), st))
            if self.scan_events is not None:
                e1.record()
                self.scan_events.append((e0, e1))
            if credit:
                # the hi-only part of eps is a worst case (2^-8); this query's actual |q/|q| - hi| is known: hand the
                # difference to the certificate by lowering the query's thresholds (tt_certificate_credit)
                check(L.tt_certificate_credit(ptr(w["cand_thresh"]), b, n_lists, ptr(w["rho"]), EPS_HI_ONLY, st))
            self._stage2(q, b, w, n_cand, n_lists, k, xchg, cert, am, eps)
        # cosine: proven iff margin > eps;  L2: the bound already contains eps, proven iff margin > 0
        return SearchResult(w["keys"], w["scores"], w["ids"], w["margin"], 0.0 if l2 else eps, bool(hi_only))

    def search_exact(self, q: torch.Tensor, k: int, out: Optional[dict] = None, rows: Optional[tuple] = None,
                     row_filter=None) -> SearchResult:
        """fp64 scoring of every row (CUDA cores): certificate-failure fallback, the chroma_l2_exp path and the
        on-GPU secondary oracle.  ``rows = (lo, hi)`` restricts the scan to that local row range (one segment)."""
        q = self._check_queries(q)
        b = int(q.shape[0])
        lo, hi = rows if rows is not None else (0, self.n_rows)
        dev = self.device
        if out is not None:
            keys, scores, ids = out["keys"], out["scores"], out["ids"]
        else:
            keys = torch.empty((b, k), dtype=torch.float32, device=dev)
            scores = torch.empty((b, k), dtype=torch.float32, device=dev)
            ids = torch.empty((b, k), dtype=torch.int64, device=dev)
        src = self.master if self.master is not None else self.corpus
        with torch.cuda.device(dev):
            nbytes = int(self.lib.tt_scan_exact_workspace_bytes(dev.index or 0, b, k))
            ws = torch.empty(max(1, nbytes), dtype=torch.uint8, device=dev)
            gate = ptr(row_filter.inv_norm[lo:hi]) if row_filter is not None else None
            check(self.lib.tt_scan_exact_f64_gated(ptr(src[lo:hi]), _lib.DTYPE_F32 if self.master is not None else _lib.DTYPE_BF16,
                                                   hi - lo, self.dim, self._row_stride(src), self.id_base + lo, ptr(q), b, k,
                                                   self.score_mode, gate, ptr(keys), ptr(scores), ptr(ids), ptr(ws), ws.numel(),
                                                   self._stream()))
        return SearchResult(keys, scores, ids, None)

    def search_certified(self, q: torch.Tensor, k: int, row_filter=None) -> SearchResult:
        """``search`` + certificate check (one device->host read of the margins); queries that are not
        proven exact climb the repair ladder.  Synchronises the current stream."""
        q = self._check_queries(q)
        r = self.search(q, k, row_filter=row_filter)
        bad = torch.nonzero(~(r.margin > r.eps)).flatten()  # NaN-safe
        if bad.numel():
            self._repair(q, k, r, bad, hi_lo_first=r.hi_only, row_filter=row_filter)
        return r

    def _rescan(self, q, k, r: SearchResult, bad: torch.Tensor, kprime: Optional[int], row_filter=None) -> torch.Tensor:
        """Re-run the queries ``bad`` hi+lo (optionally with deeper per-CTA shortlists), copy the ones that are now proven
        into ``r`` and return the indices still unproven.  The sub-batch is padded to a power of two (repeating its
        last query) so that a long-lived service keeps a handful of repair workspaces, not one per failure count."""
        n_bad = int(bad.numel())
        n_pad = 1 << max(0, n_bad - 1).bit_length()
        sel = bad if n_pad == n_bad else torch.cat([bad, bad[-1:].expand(n_pad - n_bad)])
        sub = q.index_select(0, sel)
        r2 = self.search(sub, k, out=dict(self._buffers(n_pad, k, slot=-1, hi_only=False, kprime=kprime)), hi_only=False,
                         row_filter=row_filter)
        ok = (r2.margin > r2.eps)[:n_bad]
        good = bad[ok]
        if good.numel():
            r.keys.index_copy_(0, good, r2.keys[:n_bad][ok])
            r.scores.index_copy_(0, good, r2.scores[:n_bad][ok])
            r.ids.index_copy_(0, good, r2.ids[:n_bad][ok])
        return bad[~ok]

    def _repair(self, q, k, r: SearchResult, bad: torch.Tensor, hi_lo_first: bool, row_filter=None) -> None:
        """Queries whose certificate failed climb a ladder: (a hi-only batch first gets the tighter hi+lo scan,) then a
        hi+lo re-scan with the deepest per-CTA shortlists the kernel has (K' = 128: many near-ties of the k-th score
        inside one CTA's share of the corpus are what defeats a short list), then the exact fp64 scan."""
        if hi_lo_first and bad.numel():
            self.retries += int(bad.numel())
            bad = self._rescan(q, k, r, bad, None, row_filter)
        deep = int(self.lib.tt_scan_max_kprime())
        if bad.numel() and self.kprime < deep and not os.environ.get("TT_NO_DEEP_RUNG"):
            self.deep_rescans += int(bad.numel())
            bad = self._rescan(q, k, r, bad, deep, row_filter)
        if bad.numel():
            self.fallbacks += int(bad.numel())
            ex = self.search_exact(q.index_select(0, bad), k, row_filter=row_filter)
            r.keys.index_copy_(0, bad, ex.keys)
            r.scores.index_copy_(0, bad, ex.scores)
            r.ids.index_copy_(0, bad, ex.ids)

    # ------------------------------------------------------------------ stage 3
    def automerge(self, ids: torch.Tensor, scores: torch.Tensor, ratio_thresh: float = 0.5, max_rounds: int = 64,
                  out: Optional[MergeResult] = None) -> MergeResult:
        if self.tree is None:
            raise RuntimeError("this index has no node tree: auto-merge is not available")
        b, k = int(ids.shape[0]), int(ids.shape[1])
        max_out = max(2 * k, 1)
        if out is None:
            out = MergeResult(torch.empty((b, max_out), dtype=torch.int64, device=self.device),
                              torch.empty((b, max_out), dtype=torch.float64, device=self.device),
                              torch.empty((b,), dtype=torch.int32, device=self.device))
        with self._on_device():
            check(self.lib.tt_automerge(ptr(ids), ptr(scores), b, k, ptr(self.parent_of), ptr(self.child_count),
                                        ptr(self.prev_id), ptr(self.next_id), self.n_nodes, float(ratio_thresh),
                                        int(max_rounds), ptr(out.ids), ptr(out.scores), ptr(out.lens),
                                        int(out.ids.shape[1]), self._stream()))
        return out

    def step_graph(self, b: int, k: int, ratio_thresh: float = 0.5, lane: int = 0, merged: bool = True) -> StepGraph:
        """The device pipeline of one (batch, k) shape as ONE CUDA graph over device-resident buffers: prepare ->
        stage 1 -> re-score + select + auto-merge (3 kernel nodes for batch <= 32).  ``lane`` picks an independent
        workspace set, so that graphs of different lanes may be in flight on different streams at once."""
        key = ("step", b, k, float(ratio_thresh), lane, merged)
        g = self._ws.get(key)
        if g is None:
            dev = self.device
            q = torch.zeros((b, self.dim), dtype=torch.float32, device=dev)
            w = dict(self._buffers(b, k, slot=("step", lane)))
            w["margin"] = torch.empty((b,), dtype=torch.float32, device=dev)
            mo = None
            if merged:
                mo = MergeResult(torch.empty((b, max(2 * k, 1)), dtype=torch.int64, device=dev),
                                 torch.empty((b, max(2 * k, 1)), dtype=torch.float64, device=dev),
                                 torch.empty((b,), dtype=torch.int32, device=dev))
            am = self._am_args(ratio_thresh, mo) if merged else None
            with self._on_device():
                graph, r = capture_on_side_stream(dev, lambda: self.search(q, k, out=w, am=am))
            g = self._ws[key] = StepGraph(graph, q, r, mo, r.eps)
        return g

    # ------------------------------------------------------------------ whole path, host in / host out
    def _record(self, b: int, k: int, merged: bool, extra_f32: int = 0, lane: int = 0):
        """One contiguous result record per query batch, so the whole answer (and the certificate margins)
        comes back in ONE device->host copy into pinned memory:
        ``[ lens i32 B | margin f32 B | ids i64 B*w | scores (f64 merged / f32 leaves) B*w | extra f32 ]``
        (``extra``: the row-sharded path's margins of every rank, [world, B])."""
        key = ("rec", b, k, merged, extra_f32, lane)
        r = self._ws.get(key)
        if r is None:
            w = max(2 * k, 1) if merged else k
            ssz = 8 if merged else 4
            off_ids = 8 * b
            off_sc = off_ids + 8 * b * w
            off_ex = off_sc + ssz * b * w
            total = off_ex + 4 * extra_f32

            def views(base):
                return {"lens": base[0:4 * b].view(torch.int32), "margin": base[4 * b:8 * b].view(torch.float32),
                        "ids": base[off_ids:off_sc].view(torch.int64).view(b, w),
                        "scores": base[off_sc:off_ex].view(torch.float64 if merged else torch.float32).view(b, w),
                        "extra": base[off_ex:total].view(torch.float32)}

            dev = torch.zeros(total, dtype=torch.uint8, device=self.device)
            host = torch.zeros(total, dtype=torch.uint8).pin_memory()
            hv = views(host)
            r = self._ws[key] = {"dev": dev, "host": host, "d": views(dev), "h": hv, "bytes": total,
                                 "hn": {name: t.numpy() for name, t in hv.items()},  # numpy views of the pinned record
                                 "event": torch.cuda.Event()}
        return r

    def _pipeline_graph(self, b: int, k: int, ratio_thresh: float, merged: bool, lane: int = 0):
        """The whole device pipeline of one ``retrieve_host`` shape as ONE CUDA graph: H2D of the queries from a pinned
        staging buffer -> prepare -> stage 1 -> re-score/select -> auto-merge -> D2H of the result record.  Captured
        after ``GRAPH_AFTER`` eager calls of the shape; a replay costs one launch instead of ~10 (the kernels, their
        arguments and the buffers are the same ones the eager path uses).  Returns None while the shape is still
        warming up, or for good if capture is not possible (the eager path then keeps serving)."""
        key = ("graph", b, k, float(ratio_thresh), merged, lane)
        g = self._ws.get(key)
        if g is None:
            g = self._ws[key] = {"calls": 0, "graph": None, "dead": bool(os.environ.get("TT_NO_GRAPH"))}
        if g["graph"] is not None or g["dead"]:
            return g if g["graph"] is not None else None
        g["calls"] += 1
        if g["calls"] <= GRAPH_AFTER:
            return None
        if not _CAPTURE_LOCK.acquire(blocking=False):  # someone else is capturing: stay eager this time, try again later
            return None
        try:
            vb = self._result_rows(b)
            rec = self._record(vb, k, merged, lane=lane)
            d = rec["d"]
            w = dict(self._buffers(vb, k, slot=self._host_slot(lane)))
            w["margin"] = d["margin"]
            if not merged:
                w["ids"], w["scores"] = d["ids"], d["scores"]
            q_pin = torch.zeros((b, self.dim), dtype=torch.float32).pin_memory()
            q_dev = torch.zeros((b, self.dim), dtype=torch.float32, device=self.device)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            graph = torch.cuda.CUDAGraph()
            with capturing(graph, side):
                q_dev.copy_(q_pin, non_blocking=True)
                am = self._am_args(ratio_thresh, MergeResult(d["ids"], d["scores"], d["lens"])) if merged else None
                r = self.search(q_dev, k, out=w, am=am)  # prepare -> scan -> re-score + select + auto-merge: 3 kernels
                rec["host"].copy_(rec["dev"], non_blocking=True)
            torch.cuda.current_stream(self.device).wait_stream(side)
            g.update(graph=graph, q_pin=q_pin, q_pin_np=q_pin.numpy(), q_dev=q_dev, result=r, rec=rec)
            return g
        except Exception as exc:  # capture refused (driver, another thread capturing, ...): keep the eager path
            import warnings

            g["failures"] = g.get("failures", 0) + 1
            g["dead"] = g["failures"] >= 3  # e.g. another thread synchronised the device mid-capture: worth another try
            if g["dead"]:
                warnings.warn(f"tensor_truth_b200: CUDA-graph capture of the retrieve pipeline failed ({exc}); staying eager")
            try:
                torch.cuda.synchronize(self.device)
            except Exception:
                pass
            return None
        finally:
            _CAPTURE_LOCK.release()

    @staticmethod
    def _host_slot(lane: int):
        """Workspace slot of a host lane (lane 0 shares slot 0 with plain ``search`` callers, as it always has)."""
        return 0 if lane == 0 else ("host-lane", lane)

    def retrieve_host(self, q_host: torch.Tensor, k: int, ratio_thresh: float = 0.5, merge: bool = True, row_filter=None):
        """Query embeddings in host memory -> merged ``(ids, scores, lens)`` in host memory (numpy).
        The H2D copy of the queries and the single D2H read of the result record are part of the call;
        the certificate is checked on the host from the margins in that record.  ``row_filter``: a metadata filter
        resolved by ``row_filter()`` (such calls take the eager pipeline: the captured graph reads the ungated norms)."""
        if int(q_host.shape[0]) > MAX_HOST_BATCH:  # bound the workspaces: large batches go through in slices
            parts = [self.retrieve_host(q_host[i:i + MAX_HOST_BATCH], k, ratio_thresh, merge, row_filter)
                     for i in range(0, int(q_host.shape[0]), MAX_HOST_BATCH)]
            return tuple(np.concatenate([p[j] for p in parts], axis=0) for j in range(3))
        n = int(q_host.shape[0])
        if not (self.COALESCE and self._coalesce and 0 < n <= COALESCE_MAX and q_host.dim() == 2 and q_host.shape[1] == self.dim
                and q_host.dtype == torch.float32 and q_host.device.type == "cpu"):
            with self._host_lane() as lane:
                return self._retrieve_host_lane(lane, q_host, k, ratio_thresh, merge, row_filter)
        # Callers that arrive while every lane is busy are COALESCED: whoever gets the next free lane takes all compatible
        # waiting requests along as one batch (the scan costs the same for 1 ... 8 queries), the others find their answer
        # ready.  One caller alone leads a batch of itself; two callers share the two lanes; from three on, batches form.
        req = _HostRequest(q_host, n, (int(k), float(ratio_thresh), bool(merge), id(row_filter) if row_filter is not None else 0))
        with self._lane_cv:
            self._pending.append(req)
        batch = None
        with self._host_lane(req) as lane:
            if lane >= 0:
                with self._lane_cv:
                    if not req.taken:  # (it may have been taken between getting the lane and getting here)
                        batch, total, rest = [req], n, []
                        req.taken = True
                        for r in self._pending:
                            if r is req:
                                continue
                            if r.key == req.key and total + r.n <= COALESCE_MAX:
                                r.taken = True
                                batch.append(r)
                                total += r.n
                            else:
                                rest.append(r)
                        self._pending = rest
                        if len(batch) > 1:
                            self._lane_cv.notify_all()  # the callers taken along stop queueing for a lane
            if batch is not None:
                try:
                    if len(batch) == 1:
                        req.result = self._retrieve_host_lane(lane, q_host, k, ratio_thresh, merge, row_filter)
                    else:
                        # padded to a power of two (repeating the last query): four batch shapes -- and captured graphs --
                        # per lane instead of eight, at no cost (the pass is as long for 8 queries as for 5)
                        parts = [r.q for r in batch]
                        pad = (1 << (total - 1).bit_length()) - total
                        if pad:
                            parts.append(parts[-1][-1:].expand(pad, -1))
                        ids, scores, lens = self._retrieve_host_lane(lane, torch.cat(parts), k, ratio_thresh, merge, row_filter)
                        a = 0
                        for r in batch:
                            r.result = (ids[a:a + r.n], scores[a:a + r.n], lens[a:a + r.n])
                            a += r.n
                        self.coalesced += len(batch) - 1
                except BaseException as exc:  # noqa: BLE001 -- every caller of the batch sees what its leader saw
                    for r in batch:
                        r.error = exc
                finally:
                    for r in batch[1:]:
                        r.done.set()
        if batch is None:
            req.done.wait()
        if req.error is not None:
            raise req.error
        return req.result

    def _retrieve_host_lane(self, lane: int, q_host: torch.Tensor, k: int, ratio_thresh: float, merge: bool, row_filter):
        """``retrieve_host`` on a host lane the caller holds."""
        merged = bool(merge and self.tree is not None)
        b = int(q_host.shape[0])
        g = (self._pipeline_graph(b, k, ratio_thresh, merged, lane)
             if (row_filter is None and q_host.dim() == 2 and q_host.shape[1] == self.dim) else None)
        if g is not None:
            if q_host.dtype == torch.float32 and q_host.device.type == "cpu":
                g["q_pin_np"][...] = q_host.numpy()  # the common case: a plain memory copy into the pinned staging buffer
            else:
                g["q_pin"].copy_(q_host)
            q, r, rec = g["q_dev"], g["result"], g["rec"]
            d, h = rec["d"], rec["hn"]
            with self._on_device():
                g["graph"].replay()
        else:
            q = q_host.to(self.device, torch.float32, non_blocking=True)
            q = self._check_queries(q)
            vb = self._result_rows(b)
            rec = self._record(vb, k, merged, lane=lane)
            d, h = rec["d"], rec["hn"]
            w = dict(self._buffers(vb, k, slot=self._host_slot(lane)))
            w["margin"] = d["margin"]
            if not merged:
                w["ids"], w["scores"] = d["ids"], d["scores"]
            am = self._am_args(ratio_thresh, MergeResult(d["ids"], d["scores"], d["lens"])) if merged else None
            r = self.search(q, k, out=w, am=am, row_filter=row_filter)
            rec["host"].copy_(rec["dev"], non_blocking=True)
        check(self.lib.tt_stream_synchronize(self._stream()))  # one GIL-releasing call (an event record + wait through
        _lib.check_status(self._dev_index)                     # PyTorch costs ~15 us of Python per query)
        bad = np.nonzero(~(h["margin"] > r.eps))[0]
        if bad.size:  # not proven exact: re-run those queries (tighter scan, then the exact fp64 scan)
            with self._repair_lock:
                self._repair(q, k, r, torch.from_numpy(bad).to(self.device), hi_lo_first=r.hi_only, row_filter=row_filter)
                if merged:
                    self.automerge(r.ids, r.scores, ratio_thresh, out=MergeResult(d["ids"], d["scores"], d["lens"]))
                rec["host"].copy_(rec["dev"], non_blocking=True)
                check(self.lib.tt_stream_synchronize(self._stream()))
            _lib.check_status(self._dev_index)
        ids, scores = h["ids"].copy(), h["scores"].astype(np.float64)
        lens = h["lens"].copy() if merged else (ids >= 0).sum(axis=1).astype(np.int32)
        return ids, scores, lens

    def close(self) -> None:
        """Drop device memory (``RAGService.clear`` -> ``MultiIndexRetriever.clear_cache`` path, rag_service.py:720)."""
        self._ws.clear()  # workspaces, result records, captured graphs and their pinned staging buffers
        self.corpus = self.master = self.inv_norm = None
        self.parent_of = self.child_count = self.prev_id = self.next_id = None
