#!/usr/bin/env python
"""Benchmark of the retrieval hot path (scan -> exact top-k -> auto-merge) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU arm (oracle port, all host threads)
    torchrun --nproc-per-node N ... bench.py --gpus N ...    # N > 1: one rank per GPU, corpus row-sharded

Workload (BASELINE.json configs[1]): exact cosine top-10 + auto-merge over 10,000,000 x 1024 bf16 leaf
embeddings with a 3-level node tree, batch-1 queries (the headline, HBM-bound); batch-64 is reported in
`batch64`.  A step = one query batch through the whole device pipeline.  N > 1 shards the SAME corpus
by rows over the ranks ("scaling": "strong"): local exact top-k -> one NCCL all-gather -> k-way merge
-> auto-merge on the merged list.  The corpus streamed per step (20.5 GB / N per GPU) is far larger than
the 126 MB L2, so no explicit flush is needed between steps.

Prints ONE JSON line (rank 0).
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "queries/sec exact top-10 (+auto-merge) over 10M x 1024 chunks, batch-1"
UNIT = "queries/s"
N_ROWS = 10_000_000
DIM = 1024
TOP_K = 10
LEVELS = 3
SEED = 1234
QUERY_POOL = 64


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_tensor_peak():
    """Dense bf16 TFLOP/s for a kernel timed inside a long step (the sustained figure)."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    except Exception:
        return 1400.0, "fallback (B200_PROFILING.md ~1.4 PFLOP/s sustained)"


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clocks and throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index: int, period_s: float = 0.001):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            # the first query of each kind can take tens of ms (driver-side lazy set-up): pay that here, not in the timed region
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            (getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons)(self.h)
        except Exception:
            self.nv = None

    _NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
              0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
              0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self._NAMES.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------- workload
def physical_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


def build_shard(n_rows, rank, world, device, levels=None):
    from tensor_truth_b200.sharded import shard_bounds
    from tensor_truth_b200.synth import SynthCorpus

    sc = SynthCorpus(n_rows, DIM, levels if levels is not None else LEVELS, SEED, device=device)
    lo, hi = shard_bounds(n_rows, world, rank)
    corpus, inv = sc.rows(lo, hi)
    return sc, corpus, inv, lo, hi


def make_queries(sc, corpus, lo, hi, n_q, world, first=0, lookup=None):
    import torch.distributed as dist

    if lookup is None:
        def lookup(t):
            return corpus[t - lo] if lo <= t < hi else None

    q = sc.queries(n_q, first=first, lookup=lookup)
    if world > 1:
        qd = q.cuda()
        dist.all_reduce(qd)  # every target row is owned by exactly one rank; the other ranks contribute zeros
        q = qd.cpu()
    return sc.finish_queries(q)


def cpu_arm(bits, inv_norm, tree, queries, k, n_steps, n_warm, budget_s=25.0):
    """The CPU restatement in its fast mode on every host thread (oracle/c/oracle_fast.c + oracle/automerge.py):
    one step = one batch-1 query over the rows given + auto-merge.  Returns (seconds per step, steps run)."""
    import oracle
    from oracle import cport

    def step(i):
        q = queries[i % len(queries)][None, :]
        ids, sc = cport.fast_scan_topk(bits, inv_norm, q, k)
        pairs = [(int(o), float(s)) for o, s in zip(ids[0], sc[0]) if o >= 0]
        return oracle.auto_merge(pairs, tree.parent_of, tree.child_count, tree.prev_id, tree.next_id)

    for i in range(n_warm):
        step(i)
    t0 = time.perf_counter()
    done = 0
    for i in range(n_steps):
        step(n_warm + i)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    return (time.perf_counter() - t0) / done, done


def host_can_hold(n_bytes: int) -> bool:
    """Whether the CPU arm may keep ``n_bytes`` of corpus in host memory (with the same again as slack)."""
    try:
        import psutil

        return psutil.virtual_memory().available > 2 * n_bytes + (8 << 30)
    except Exception:
        return False


def to_host_bits(corpus: torch.Tensor) -> np.ndarray:
    """bf16 device rows -> uint16 host array, copied in 1 GB slices through a pinned bounce buffer."""
    n, d = int(corpus.shape[0]), int(corpus.shape[1])
    out = np.empty((n, d), dtype=np.uint16)
    step = max(1, (1 << 30) // (2 * d))
    pin = torch.empty((step, d), dtype=torch.int16).pin_memory()
    view = corpus.view(torch.int16)
    for a in range(0, n, step):
        m = min(step, n - a)
        pin[:m].copy_(view[a:a + m])
        out[a:a + m] = pin[:m].numpy().view(np.uint16)
    return out


# --------------------------------------------------------------------------- this repo's arm
class Bench:
    """One corpus resident on this rank + the timing loops over it.  ``run_b200`` builds one per section."""

    def __init__(self, ctx, n_rows, levels, k, kprime, variant, n_pool=QUERY_POOL, master_f32=False):
        from tensor_truth_b200.index import DeviceIndex
        from tensor_truth_b200.sharded import ShardedIndex

        self.ctx = ctx
        self.k = k
        self.n_rows = n_rows
        self.sc, corpus, inv, self.lo, self.hi = build_shard(n_rows, ctx.rank, ctx.world, ctx.device, levels)
        self.queries = make_queries(self.sc, corpus, self.lo, self.hi, n_pool, ctx.world).to(ctx.device)
        self.corpus_bf16, self.inv = corpus, inv
        if master_f32:
            # an fp32 store (what a real bge-m3 index holds, indexing/builder.py:437-442): the canonical bf16 rows plus
            # low-order mantissa bits the bf16 shadow cannot represent (deterministic, relative size ~2^-9)
            master = torch.empty((self.hi - self.lo, DIM), dtype=torch.float32, device=ctx.device)
            step = 1 << 20
            for a in range(0, self.hi - self.lo, step):
                c = corpus[a:a + step].float()
                g = torch.Generator(device=ctx.device)
                g.manual_seed(SEED * 31 + (self.lo + a) // step)
                master[a:a + step] = c * (1.0 + (2.0 ** -9) * (torch.rand(c.shape, generator=g, device=ctx.device) - 0.5))
            self.idx = DeviceIndex(master, self.sc.tree, id_base=self.lo, device=ctx.device, kprime=kprime, variant=variant)
            del master, corpus, inv
            self.corpus_bf16, self.inv = self.idx.corpus, self.idx.inv_norm  # the shadow the index derived from the master
            torch.cuda.empty_cache()
        else:
            self.idx = DeviceIndex(corpus, self.sc.tree, inv_norm=inv, id_base=self.lo, device=ctx.device, kprime=kprime,
                                   variant=variant)
        self.sharded = ShardedIndex(self.idx) if ctx.world > 1 else None
        self.rows_local = self.hi - self.lo

    @classmethod
    def head_of(cls, parent: "Bench", n_head: int, k: int, variant):
        """A Bench over the first ``n_head`` rows of every rank's shard of ``parent`` (shared memory, own index): how the
        C4 section gets 6.25M rows per GPU out of the C3 corpus without generating another one."""
        from tensor_truth_b200.index import DeviceIndex
        from tensor_truth_b200.sharded import ShardedIndex

        ctx = parent.ctx
        w = cls.__new__(cls)
        w.ctx, w.k, w.sc = ctx, k, parent.sc
        w.lo, w.hi, w.rows_local, w.n_rows = parent.lo, parent.lo + n_head, n_head, n_head * ctx.world
        w.corpus_bf16, w.inv = parent.corpus_bf16[:n_head], parent.inv[:n_head]
        w.idx = DeviceIndex(w.corpus_bf16, parent.sc.tree, inv_norm=w.inv, id_base=parent.lo, device=ctx.device, variant=variant)
        w.sharded = ShardedIndex(w.idx) if ctx.world > 1 else None
        w.queries = parent.queries
        # a query's target row t of the parent corpus is answered by its owner with row (t - lo) % n_head of the head
        p_lo, p_hi, head = parent.lo, parent.hi, w.corpus_bf16
        w.query_lookup = lambda t: head[(t - p_lo) % n_head] if p_lo <= t < p_hi else None
        return w

    query_lookup = None

    def close(self):
        if self.sharded is not None:
            self.sharded.close()
        else:
            self.idx.close()
        self.corpus_bf16 = self.inv = self.queries = None
        torch.cuda.empty_cache()

    # ---- one step, eager (serial loop: per-kernel events) and as a graph (pipelined loop)
    def step_graph(self, batch, k, lane):
        if self.sharded is not None and self.sharded.transport == "peer":
            return self.sharded.step_graph(batch, k, 0.5, lane)
        if self.sharded is None:
            return self.idx.step_graph(batch, k, 0.5, lane)
        return None  # NCCL transport: eager steps

    def timed(self, batch: int, steps: int, warm: int, sample_clocks: bool, depth: int, k: int = 0, pool=None,
              serial_graph: bool = False):
        """K steps of the device pipeline, queries resident in HBM.
        depth = 1: strictly serial on one stream, eager launches with CUDA events around every stage-1 launch (the
        roofline numbers; the host stays ahead of the GPU, so the events see device time only).
        depth = 2: the throughput configuration -- every lane (stream) replays the CUDA graph of the whole step
        (prepare -> scan -> re-score+select(+push) -> (flag-wait+merge+)auto-merge), steps alternate between the lanes, so
        step i+1's scan overlaps step i's tail.
        depth = 1 with ``serial_graph``: one lane's graph replayed back to back on one stream -- the serial step without
        the host's launch gaps (what the device alone needs for scan + tail).
        All loops start behind a device-side rendezvous of all ranks."""
        from tensor_truth_b200.index import MergeResult

        ctx, idx, sharded = self.ctx, self.idx, self.sharded
        device = ctx.device
        k = k or self.k
        pool = self.queries if pool is None else pool
        n_pool = max(1, int(pool.shape[0]) // batch)
        margins = torch.full((steps + warm, batch), float("inf"), dtype=torch.float32, device=device)
        streams = [torch.cuda.Stream(device) for _ in range(depth)]
        eps = [idx.eps]
        graphs = [self.step_graph(batch, k, s) for s in range(depth)] if (depth > 1 or serial_graph) else [None]
        use_graph = graphs[0] is not None
        launches = [0]
        if not use_graph:
            bufs = [idx._buffers(batch, k, slot=s) for s in range(depth)]
            mouts = [MergeResult(torch.empty((batch, 2 * k), dtype=torch.int64, device=device),
                                 torch.empty((batch, 2 * k), dtype=torch.float64, device=device),
                                 torch.empty((batch,), dtype=torch.int32, device=device)) for _ in range(depth)]

        def one(i):
            s = i % depth
            q = pool[(i % n_pool) * batch:(i % n_pool) * batch + batch]
            with torch.cuda.stream(streams[s]):
                if use_graph:
                    g = graphs[s]
                    g.q.copy_(q, non_blocking=True)
                    g.replay()
                    margins[i].copy_(g.result.margin, non_blocking=True)
                    eps[0] = g.eps
                    return g.merged
                if sharded is None:
                    w = dict(bufs[s])
                    w["margin"] = margins[i]
                    r = idx.search(q, k, out=w, am=idx._am_args(0.5, mouts[s]))
                    eps[0] = r.eps
                    return mouts[s]
                sharded.search(q, k, margins=margins[i], slot=s, merged_out=mouts[s])
                eps[0] = sharded.last.eps
                return mouts[s]

        cur = torch.cuda.current_stream()
        for st in streams:
            st.wait_stream(cur)
        for i in range(warm):
            one(i)
        for st in streams:
            cur.wait_stream(st)
        ctx.barrier()
        idx.scan_events = [] if (depth == 1 and not use_graph) else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(physical_index(ctx.local_rank)) if sample_clocks else None
        if sampler:
            sampler.__enter__()
        if sharded is not None:
            sharded.barrier(batch, k)  # device-side rendezvous: rank skew stays outside the timed region
        e0.record()
        for st in streams:
            st.wait_stream(cur)
        for i in range(steps):
            last = one(warm + i)
        for st in streams:
            cur.wait_stream(st)
        e1.record()
        ctx.barrier()
        if sampler:
            sampler.__exit__()
        ms = ctx.max_over_ranks(e0.elapsed_time(e1))
        ev = idx.scan_events
        idx.scan_events = None
        scan_ms, n_scan_launches = None, None
        if ev:
            scan_ms = ctx.max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in ev])))
            n_scan_launches = len(ev) // steps
        from tensor_truth_b200 import _lib

        _lib.check_status(idx._dev_index)
        mt = margins[warm:]
        bad = int((~(mt > eps[0])).sum().item())
        return {"ms": ms, "scan_ms": scan_ms, "scan_calls_per_step": n_scan_launches, "bad": bad, "last": last,
                "clocks": sampler.summary() if sampler else None, "min_margin": float(mt.min().item()), "eps": eps[0],
                "graph": use_graph}

    def hbm_section(self, steps, warm, peak, peak_src, batch=1, sample_clocks=False, label=""):
        """Serial loop (roofline of the stage-1 kernel) + pipelined loop (throughput) at one batch size."""
        ser = self.timed(batch, steps, warm, False, depth=1)
        sg = self.timed(batch, steps, warm, False, depth=1, serial_graph=True)
        pip = self.timed(batch, steps, warm, sample_clocks, depth=2)
        local_bytes = float(self.rows_local) * DIM * 2
        achieved = local_bytes / (ser["scan_ms"] / 1e3) / 1e9
        return {
            "value": steps * batch / (pip["ms"] / 1e3), "unit": UNIT, "ms_per_step": pip["ms"] / steps, "steps": steps,
            "warmup": warm, "batch": batch, "k": self.k, "rows_total": self.n_rows, "rows_per_gpu": self.rows_local,
            "serial": {"value": steps * batch / (ser["ms"] / 1e3), "ms_per_step": ser["ms"] / steps},
            "serial_graph": ({"value": steps * batch / (sg["ms"] / 1e3), "ms_per_step": sg["ms"] / steps,
                              "note": "same K steps, one lane's CUDA graph replayed back to back on one stream: no overlap between "
                                      "steps and no host launch gaps"} if sg["graph"] else None),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "kernel": "scan_tc_kernel", "bytes_per_launch": local_bytes,
                         "kernel_ms": ser["scan_ms"], "peak_source": peak_src,
                         "step_share": ser["scan_ms"] * ser["scan_calls_per_step"] / (ser["ms"] / steps),
                         "step_share_graph": (ser["scan_ms"] * ser["scan_calls_per_step"] / (sg["ms"] / steps)) if sg["graph"] else None,
                         "measured_in": "serial loop, CUDA events around each stage-1 launch"},
            "hbm_frac_of_step": local_bytes / (pip["ms"] / steps / 1e3) / 1e9 / peak,
            "certificate_failures": ser["bad"] + pip["bad"], "min_margin": pip["min_margin"], "eps": pip["eps"],
            "pipelined_with_cuda_graphs": pip["graph"],
        }, ser, pip

    def wide_section(self, bw, kw, steps):
        """BASELINE configs[3]'s shape: `bw` concurrent queries, top-`kw` + auto-merge, through the GEMM-shaped stage 1."""
        ctx, idx = self.ctx, self.idx
        qw = make_queries(self.sc, self.corpus_bf16, self.lo, self.hi, bw, ctx.world, lookup=self.query_lookup).to(ctx.device)
        tpeak, tpeak_src = measured_tensor_peak()
        serw = self.timed(bw, steps, 1, False, depth=1, k=kw, pool=qw)
        flops = 2.0 * bw * float(self.rows_local) * DIM
        tfl = flops / (serw["scan_ms"] / 1e3) / 1e12
        # spot check: the first 4 queries against the exact fp64 scan of the local shard
        rw = idx.search(qw, kw, out=dict(idx._buffers(bw, kw, slot=0)))
        exw = idx.search_exact(qw[:4], kw)
        wide_ok = bool(torch.equal(rw.ids[:4], exw.ids) and torch.equal(rw.scores[:4], exw.scores))
        out = {"value": steps * bw / (serw["ms"] / 1e3), "unit": UNIT, "batch": bw, "k": kw,
               "ms_per_step": serw["ms"] / steps, "steps": steps, "rows_total": self.n_rows, "rows_per_gpu": self.rows_local,
               "stage1_ms_per_step": serw["scan_ms"], "gemm_path": bool(idx._use_gemm(bw)),
               "roofline": {"bound": "tensor", "achieved": tfl, "peak": tpeak, "unit": "TFLOP/s", "frac": tfl / tpeak,
                            "traffic": None, "kernel": "scan_gemm_kernel (+ gemm_cut_kernel between phases)",
                            "flops_per_step": flops, "peak_source": tpeak_src,
                            "measured_in": "CUDA events around the whole stage 1 of each step (all phases and cuts)"},
               "certificate_failures": serw["bad"], "min_margin": serw["min_margin"], "eps": serw["eps"],
               "parity_vs_gpu_exact_scan_local_shard": wide_ok,
               "mode": "hi-only bf16 queries, 256 x 256 tcgen05 pair tiles, data-driven thresholds in phases"}
        del qw, rw, exw
        idx._ws = {kk: v for kk, v in idx._ws.items() if not (isinstance(kk, tuple) and kk and kk[0] == bw)}
        if self.sharded is not None:
            self.sharded._peers.pop((bw, kw), None)
            self.sharded._out.clear()
        torch.cuda.empty_cache()
        return out

    def retriever(self):
        from tensor_truth_b200.retriever import B200AutoMergingRetriever, B200VectorIndexRetriever

        base_r = B200VectorIndexRetriever(self.idx if self.sharded is None else self.sharded, similarity_top_k=self.k)
        return B200AutoMergingRetriever(base_r, None)  # over a ShardedIndex every rank makes the same call (SPMD)

    def e2e_section(self, steps, q_host=None, warm=8, many_callers=False):
        """End to end through the public retriever API: host query embedding in, List[NodeWithScore] out; H2D of the
        query and D2H of the result record inside the timed region (wall clock, max over ranks)."""
        from tensor_truth_b200.schema import QueryBundle

        ctx = self.ctx
        q_host = self.queries.cpu() if q_host is None else q_host
        q_lists = [row.tolist() for row in q_host]  # host input as an embed model hands it over: a Python list of floats
        am = self.retriever()
        n = len(q_lists)
        call = lambda i: am.retrieve(QueryBundle(query_str=f"q{i}", embedding=q_lists[i % n]))  # noqa: E731
        for i in range(warm):  # past GRAPH_AFTER: the one-off CUDA-graph capture of the pipeline belongs to the warm-up
            out = call(i)
        ctx.barrier()
        t0 = time.perf_counter()
        for i in range(steps):
            out = call(warm + i)
        ctx.barrier()
        e2e_s = ctx.max_over_ranks(time.perf_counter() - t0)
        two = guarded("e2e_two_callers", lambda: self.e2e_two_callers(am, q_lists, steps, warm), sys.stderr)
        eight = None
        if self.sharded is None and many_callers:  # one GPU: a serving process with many requests in flight (coalescing)
            # (long warm-up: every coalesced batch shape -- 1, 2, 4, 8 queries per lane -- captures its graph on first use)
            eight = guarded("e2e_eight_callers", lambda: self.e2e_callers(am, q_lists, max(steps, 160), 60, 8), sys.stderr)
        idx = self.idx
        d2h = idx._record(1, self.k, True, extra_f32=ctx.world if (self.sharded is not None and self.sharded.transport == "peer") else 0)["bytes"]
        return {"value": steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": DIM * 4, "d2h_bytes_per_step": d2h,
                "steps": steps, "warmup": warm,
                "api": "B200AutoMergingRetriever.retrieve(QueryBundle)" + ("" if self.sharded is None else " over ShardedIndex, every rank"),
                "retries": idx.retries, "deep_rescans": idx.deep_rescans, "fallbacks": idx.fallbacks,
                "second_rounds": getattr(self.sharded, "second_rounds", 0) if self.sharded is not None else 0,
                "nodes_returned_last": len(out), "callers": 1, "two_callers": two, "eight_callers": eight}

    def e2e_two_callers(self, am, q_lists, steps, warm):
        return self.e2e_callers(am, q_lists, steps, warm, 2)

    def e2e_callers(self, am, q_lists, steps, warm, callers):
        """The same ``retrieve()`` calls from TWO host threads (what a serving process with concurrent requests does, and
        what ``value``'s two device lanes correspond to): the index pipelines them over its host lanes, so one caller's
        exchange + D2H + Python work overlaps the other's corpus scan.  ``steps`` calls in total, wall clock, max over ranks.
        On a sharded index thread t of every rank serves lane t (``retriever.for_lane(t)``)."""
        import threading

        from tensor_truth_b200.schema import QueryBundle

        ctx, n = self.ctx, len(q_lists)
        rets = [am.for_lane(t) if self.sharded is not None else am for t in range(callers)]
        gate = threading.Barrier(callers + 1)
        errors, last = [], [None] * callers

        def worker(t):
            try:
                torch.cuda.set_device(ctx.device)
                call = lambda i: rets[t].retrieve(QueryBundle(query_str=f"q{i}", embedding=q_lists[i % n]))  # noqa: E731
                for i in range(warm):
                    call(i)
                gate.wait()   # warm-up done (graphs of both lanes captured)
                gate.wait()   # clock started
                for i in range(t, steps, callers):
                    last[t] = call(warm + i)
            except Exception as exc:  # noqa: BLE001
                errors.append(exc)
                gate.abort()

        threads = [threading.Thread(target=worker, args=(t,), daemon=True) for t in range(callers)]
        for th in threads:
            th.start()
        try:
            gate.wait()
            ctx.barrier()
            t0 = time.perf_counter()
            gate.wait()
        except threading.BrokenBarrierError:
            pass
        for th in threads:
            th.join()
        if errors:
            raise errors[0]
        ctx.barrier()
        sec = ctx.max_over_ranks(time.perf_counter() - t0)
        ref = am.retrieve(QueryBundle(query_str="check", embedding=q_lists[(warm + steps - 1) % n]))
        same = [(x.node.node_id, x.score) for x in last[(steps - 1) % callers]] == [(x.node.node_id, x.score) for x in ref]
        out = {"value": steps / sec, "unit": UNIT, "callers": callers, "steps": steps,
               "equals_single_caller_answer": bool(same),
               "note": "two host threads, each a synchronous retrieve(); pipelined over the index's host lanes"}
        if callers > 2:
            out["note"] = (f"{callers} host threads, each a synchronous retrieve(): callers that find both lanes busy are coalesced into "
                           "one batch by whoever gets the next lane (a scan pass costs the same for 1 ... 8 queries)")
            out["queries_coalesced"] = int(self.idx.coalesced)
        return out


class Ctx:
    def __init__(self):
        import torch.distributed as dist

        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.device = torch.device(f"cuda:{self.local_rank}")
        if self.world > 1:
            # NCCL writes its version banner (any NCCL_DEBUG level >= VERSION, which this image sets) and its logs to
            # stdout; rank 0 must print ONE line there
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
                os.environ["NCCL_DEBUG"] = "NONE"
            dist.init_process_group("nccl", device_id=self.device)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def guarded(name, fn, log):
    """Secondary sections never cost the headline line: an exception becomes {"error": ...}."""
    t0 = time.perf_counter()
    try:
        out = fn()
    except Exception as exc:
        import traceback

        traceback.print_exc(file=sys.stderr)
        out = {"error": f"{type(exc).__name__}: {exc}"}
        try:
            torch.cuda.synchronize()
        except Exception:
            pass
    if isinstance(out, dict):
        out["section_seconds"] = round(time.perf_counter() - t0, 2)
    print(f"[bench] section {name}: {time.perf_counter() - t0:.1f} s", file=log, flush=True)
    return out


def oracle_parity_c1(ctx, kprime, variant):
    """BASELINE configs[0] (C1: 100k rows, 64 queries, top-10 + auto-merge, 3 levels) through the SAME path the timed
    loops use -- row-sharded over all ranks when N > 1 -- checked on rank 0 against the CPU oracle (strict fp64 C
    restatement + oracle/automerge.py) on the whole corpus: ids and scores bit-equal."""
    from tensor_truth_b200.index import MergeResult

    n, nq, k = 100_000, 64, 10
    w = Bench(ctx, n, 3, k, kprime, variant, n_pool=nq)
    dev = ctx.device
    got = []
    for a in range(0, nq, 8):
        q = w.queries[a:a + 8]
        mo = MergeResult(torch.empty((8, 2 * k), dtype=torch.int64, device=dev), torch.empty((8, 2 * k), dtype=torch.float64, device=dev),
                         torch.empty((8,), dtype=torch.int32, device=dev))
        if w.sharded is not None:
            scores, ids = w.sharded.search(q, k, merged_out=mo)
        else:
            r = w.idx.search(q, k, am=w.idx._am_args(0.5, mo))
            scores, ids = r.scores, r.ids
        torch.cuda.synchronize()
        got.append((ids.cpu().numpy().copy(), scores.cpu().numpy().copy(), mo.ids.cpu().numpy().copy(),
                    mo.scores.cpu().numpy().copy(), mo.lens.cpu().numpy().copy()))
    # the whole corpus on rank 0's host: every rank generates its shard, rank 0 gathers the bits
    bits_local = w.corpus_bf16.view(torch.int16)
    if ctx.world > 1:
        parts = [torch.empty((shard_rows(n, ctx.world, r), DIM), dtype=torch.int16, device=dev) for r in range(ctx.world)]
        for r in range(ctx.world):  # broadcast each shard in turn (shards may differ in size); NCCL has no int16: send the bytes
            if r == ctx.rank:
                parts[r].copy_(bits_local)
            ctx.dist.broadcast(parts[r].view(torch.uint8), src=r)
        bits_all = torch.cat(parts, dim=0)
    else:
        bits_all = bits_local
    ok, checked = True, 0
    if ctx.rank == 0:
        import oracle
        from oracle import cport

        cport.build()
        bits = bits_all.cpu().numpy().view(np.uint16)
        qh = w.queries.cpu().numpy()
        ids_o, sc_o, _ = cport.scan_topk(bits, qh, k)
        tree = w.sc.tree
        for bi, a in enumerate(range(0, nq, 8)):
            ids, scores, mids, msc, mlens = got[bi]
            ok = ok and bool((ids == ids_o[a:a + 8]).all() and (scores == sc_o[a:a + 8]).all())
            for j in range(8):
                exp = oracle.auto_merge([(int(o), float(s)) for o, s in zip(ids_o[a + j], sc_o[a + j]) if o >= 0],
                                        tree.parent_of, tree.child_count, tree.prev_id, tree.next_id)
                gl = [(int(o), float(s)) for o, s in zip(mids[j, :mlens[j]], msc[j, :mlens[j]])]
                ok = ok and gl == exp
                checked += 1
    w.close()
    return {"ok": bool(ok), "queries_checked": checked, "rows": n, "k": k,
            "what": "C1 through the timed path (sharded over all ranks when N > 1) vs the CPU oracle on rank 0: ids, fp32 scores, merged ids, fp64 merged scores bit-equal"}


def shard_rows(n, world, r):
    from tensor_truth_b200.sharded import shard_bounds

    lo, hi = shard_bounds(n, world, r)
    return hi - lo


def hard_section(w, n_dup_per_cta, steps):
    """Queries aimed at near-duplicate rows: `n_dup_per_cta` x 148 rows of the shard are overwritten with copies of one
    base vector perturbed by ~1e-3 (relative), spread over the whole shard so that every CTA's share holds about that
    many; the query is the base vector.  More near-ties of the k-th score than a CTA's shortlist is deep defeats the
    certificate, and the repair ladder has to answer: hi+lo re-scan with K' = 128 shortlists, then the exact fp64 scan.
    Timed through retrieve_host (the path that enforces the certificate), wall clock."""
    ctx, idx = w.ctx, w.idx
    dev = ctx.device
    g = torch.Generator(device="cpu")
    g.manual_seed(SEED + 4242)
    base = torch.randn(DIM, generator=g)
    base = base / base.norm()
    n_dup_global = n_dup_per_cta * 148 * ctx.world
    stride = max(1, w.n_rows // n_dup_global)
    rows = torch.arange(0, n_dup_global, dtype=torch.int64) * stride
    rows = rows[(rows >= w.lo) & (rows < w.hi)] - w.lo
    gd = torch.Generator(device=dev)
    gd.manual_seed(SEED + 99 + ctx.rank)
    saved = idx.corpus[rows.to(dev)].clone()
    saved_inv = idx.inv_norm[rows.to(dev)].clone()
    dup = (base.to(dev)[None, :] * (1.0 + 1e-3 * torch.randn((rows.numel(), DIM), generator=gd, device=dev))).to(torch.bfloat16)
    idx.corpus[rows.to(dev)] = dup
    idx.inv_norm[rows.to(dev)] = dup.float().pow(2).sum(dim=1).clamp_min(1e-30).rsqrt()
    before = (idx.retries, idx.deep_rescans, idx.fallbacks)
    target = w.sharded if w.sharded is not None else idx
    qh = (base[None, :] * (1.0 + 1e-3 * torch.randn((8, DIM), generator=g))).float()
    qh = qh / qh.norm(dim=1, keepdim=True)
    ok = True
    for i in range(4):  # past GRAPH_AFTER: the one-off graph capture of this call shape stays outside the timed region
        target.retrieve_host(qh[i:i + 1], w.k, merge=False)
    ctx.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        ids, scores, lens = target.retrieve_host(qh[i % 8:i % 8 + 1], w.k, merge=False)
    ctx.barrier()
    sec = ctx.max_over_ranks(time.perf_counter() - t0)
    # the answer must still be the exact one: compare the last query with the exact fp64 scan of every shard
    ex = idx.search_exact(qh[(steps - 1) % 8:(steps - 1) % 8 + 1].to(dev), w.k)
    if w.sharded is None:
        ok = bool((ex.ids.cpu().numpy() == ids).all())
    out = {"value": steps / sec, "unit": UNIT, "steps": steps, "near_duplicates_per_cta": n_dup_per_cta,
           "near_duplicates_total": int(n_dup_global), "retries": idx.retries - before[0],
           "deep_rescans_kprime128": idx.deep_rescans - before[1], "exact_fp64_fallbacks": idx.fallbacks - before[2],
           "second_rounds": getattr(w.sharded, "second_rounds", 0) if w.sharded is not None else 0,
           "matches_exact_scan": ok if w.sharded is None else None}
    idx.corpus[rows.to(dev)] = saved
    idx.inv_norm[rows.to(dev)] = saved_inv
    return out


def run_b200(args):
    from tensor_truth_b200 import _lib

    ctx = Ctx()
    world, rank = ctx.world, ctx.rank
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus}")
    _lib.lib()  # fail loudly if the CUDA library is missing
    log = sys.stderr
    variant = {"auto": _lib.SCAN_AUTO, "simt": _lib.SCAN_SIMT, "tcgen05": _lib.SCAN_TCGEN05}[args.variant]
    peak, peak_src = measured_peaks()
    skip = set(x for x in args.skip.split(",") if x)
    t_start = time.perf_counter()

    # ================= headline: BASELINE configs[1] (C2), batch-1, the same rows at every N (strong scaling)
    w = Bench(ctx, args.rows, LEVELS, TOP_K, args.kprime, variant)
    print(f"[bench] corpus {args.rows} rows built in {time.perf_counter() - t_start:.1f} s", file=log, flush=True)
    head, ser1, pip1 = w.hbm_section(args.steps, args.warmup, peak, peak_src, batch=1, sample_clocks=True)
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "scan_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("rows") == w.rows_local and tj.get("batch") == 1:
            traffic, traffic_src = tj.get("dram_bytes_per_launch"), "ncu --set full capture: " + str(tj.get("source"))
    except Exception:
        pass
    head["roofline"]["traffic"] = traffic
    head["roofline"]["traffic_source"] = traffic_src or "no ncu capture for this shard size (profiles/scan_traffic.json is for the 1-GPU shard)"
    nl1 = ser1["scan_calls_per_step"]

    # ---- batch-64 (same corpus, one pass of 64 hi-only queries; tensor work rises, HBM bytes per pass do not)
    batch64 = None
    if "batch64" not in skip:
        def f64():
            steps64 = max(3, min(args.steps, 40))
            ser64 = w.timed(64, steps64, 3, False, depth=1)
            pip64 = w.timed(64, steps64, 3, False, depth=2)
            return {"value": steps64 * 64 / (pip64["ms"] / 1e3), "unit": UNIT, "ms_per_step": pip64["ms"] / steps64,
                    "steps": steps64, "serial_value": steps64 * 64 / (ser64["ms"] / 1e3), "scan_ms_per_step": ser64["scan_ms"],
                    "hbm_frac": float(w.rows_local) * DIM * 2 / (ser64["scan_ms"] / 1e3) / 1e9 / peak,
                    "certificate_failures": ser64["bad"] + pip64["bad"], "min_margin": pip64["min_margin"],
                    "eps": pip64["eps"], "mode": "hi-only bf16 queries, 64 per corpus pass",
                    "pipelined_with_cuda_graphs": pip64["graph"],
                    "kernel": "scan_gemm_kernel<64,2>" if w.idx._use_gemm(64) else "scan_tc2_kernel<64,2>"}
        batch64 = guarded("batch64", f64, log)

    # ---- wide batch on the headline corpus (BASELINE configs[3]'s shape: 16k concurrent queries, top-100)
    wide = None
    if args.wide_batch > 0 and "wide" not in skip:
        wide = guarded("wide", lambda: w.wide_section(args.wide_batch, args.wide_k, args.wide_steps), log)

    # ---- parity spot check inside the bench: the timed path vs the on-GPU exact fp64 scan of the same shard(s)
    qs = w.queries[:4]
    if w.sharded is None:
        r = w.idx.search(qs, TOP_K)
        got_ids, got_sc = r.ids.clone(), r.scores.clone()
        ex = w.idx.search_exact(qs, TOP_K)
        parity_ok = bool(torch.equal(got_ids, ex.ids) and torch.equal(got_sc, ex.scores))
    else:
        scores, ids = w.sharded.search(qs, TOP_K)
        got_ids, got_sc = ids.clone(), scores.clone()
        ex = w.idx.search_exact(qs, TOP_K)
        w.sharded.plumbing.local_search = lambda q, k, ko, io, slot=0: (ko.copy_(ex.keys), io.copy_(ex.ids))
        s2, i2 = w.sharded.plumbing.search(qs, TOP_K)
        parity_ok = bool(torch.equal(got_ids, i2) and torch.equal(got_sc, s2))
        w.sharded.plumbing.local_search = w.sharded._local_search
    torch.cuda.synchronize()

    # ---- end to end through the public retriever API
    e2e = None
    if "e2e" not in skip:
        e2e = guarded("e2e", lambda: w.e2e_section(max(5, min(args.steps, 100)), many_callers=True), log)

    # ---- CPU baseline (rank 0, N=1 only): the SAME corpus bytes on the host cores, all rows when host memory allows
    cpu = None
    if world == 1 and not args.no_cpu and "cpu" not in skip:
        def fcpu():
            from oracle import cport

            cport.build()
            cport.use_all_host_threads()
            full = host_can_hold(w.rows_local * DIM * 2) and not args.cpu_sample_rows
            sample = w.rows_local if full else min(args.cpu_sample_rows or 1_048_576, w.rows_local)
            bits = to_host_bits(w.corpus_bf16[:sample])
            inv_h = w.inv[:sample].cpu().numpy()
            sec, done = cpu_arm(bits, inv_h, w.sc.tree, w.queries.cpu().numpy(), TOP_K, 20, 2, budget_s=20.0)
            scale = sample / w.n_rows
            return {"value": (1.0 / sec) * scale, "unit": UNIT, "cores": cport.fast_threads(), "kind": "port",
                    "sample": (f"{done} batch-1 queries over " + (f"ALL {sample} rows" if full else f"the first {sample} of {w.n_rows} rows, q/s scaled by {scale:.4f}")
                               + f" ({sec * 1e3:.1f} ms each); oracle/c/oracle_fast.c fp32 AVX2 + OpenMP, auto-merge in oracle/automerge.py"),
                    "full_corpus": bool(full), "host_cpus": os.cpu_count()}
        cpu = guarded("cpu_baseline", fcpu, log)

    # ---- hard queries: near-duplicate rows defeat the certificate, the repair ladder answers (mutates rows, restores them)
    hard = None
    if "hard" not in skip:
        def fhard():
            return {"deep_rung": hard_section(w, 64, 10), "exact_fallback": hard_section(w, 200, 6),
                    "what": "queries aimed at 64 / 200 near-duplicate rows per CTA share (K' = 32 lists): K' = 128 re-scan, then exact fp64 scan"}
        hard = guarded("hard", fhard, log)
    w.close()
    del w

    # ================= fp32 store at the headline size (what a real bge-m3 index holds): bf16 shadow scanned, fp32 master re-scored
    fp32_store = None
    if "fp32" not in skip:
        def ffp32():
            wf = Bench(ctx, args.rows, LEVELS, TOP_K, args.kprime, variant, master_f32=True)
            sec, _s, _p = wf.hbm_section(max(5, min(args.steps, 40)), 3, peak, peak_src, batch=1)
            e = wf.e2e_section(max(5, min(args.steps, 40)))
            sec["e2e"] = e
            sec["store"] = "fp32 master [rows, 1024] (re-scored in fp64) + bf16 shadow (scanned); eps = 4.2e-3"
            sec["retries"], sec["deep_rescans"], sec["fallbacks"] = wf.idx.retries, wf.idx.deep_rescans, wf.idx.fallbacks
            wf.close()
            return sec
        fp32_store = guarded("fp32_store", ffp32, log)

    # ================= BASELINE configs[2] (C3): 12.5M rows per GPU (100M over 8), batch-1 -- weak scaling
    c3 = c4 = None
    if "c3" not in skip or "c4" not in skip:
        def fc3():
            rows3 = args.c3_rows_per_gpu * world
            w3 = Bench(ctx, rows3, 3, 10, 32, variant)
            out = {}
            if "c3" not in skip:
                sec, _s, _p = w3.hbm_section(max(5, min(args.steps, 40)), 3, peak, peak_src, batch=1)
                sec["scaling"] = "weak"
                sec["config"] = f"C3: {rows3} x {DIM} rows over {world} GPU(s) ({args.c3_rows_per_gpu} per GPU), batch-1, top-10 + auto-merge"
                sec["aggregate_hbm_frac"] = sec["hbm_frac_of_step"]
                sec["e2e"] = w3.e2e_section(max(5, min(args.steps, 40)))
                out["c3"] = sec
            if "c4" not in skip:
                # BASELINE configs[3] (C4): 50M rows over 8 GPUs = 6.25M per GPU, 16 384 queries, top-100.  The first
                # c4_rows_per_gpu rows of each rank's C3 shard stand for it (same generator, same bytes per GPU).
                n4 = min(args.c4_rows_per_gpu, w3.rows_local)
                w3.idx._ws.clear()
                w4 = Bench.head_of(w3, n4, args.wide_k, variant)
                sec4 = w4.wide_section(args.wide_batch, args.wide_k, args.wide_steps)
                sec4["scaling"] = "weak"
                sec4["config"] = (f"C4: {n4 * world} x {DIM} rows over {world} GPU(s) ({n4} per GPU: the head of each rank's C3 shard), "
                                  f"{args.wide_batch} queries per step, top-{args.wide_k} + auto-merge")
                out["c4"] = sec4
            w3.close()
            return out
        r34 = guarded("c3+c4", fc3, log)
        c3, c4 = r34.get("c3"), r34.get("c4")
        if "error" in r34:
            c3 = c3 or {"error": r34["error"]}

    # ================= BASELINE configs[4] (C5): 20M leaves, 4 levels, top-200, the same rows at every N
    c5 = None
    if "c5" not in skip:
        def fc5():
            w5 = Bench(ctx, args.c5_rows, 4, 200, 128, variant)
            sec, _s, _p = w5.hbm_section(max(5, min(args.steps, 20)), 3, peak, peak_src, batch=1)
            sec["scaling"] = "strong"
            sec["config"] = f"C5: {args.c5_rows} leaves, 4-level tree, top-200 + auto-merge, batch-1, over {world} GPU(s)"
            sec["e2e"] = w5.e2e_section(max(5, min(args.steps, 20)))
            # merges must actually fire on this workload: report the last merged list's length against k
            sec["merged_len_last"] = int(_p["last"].lens[0].item()) if _p.get("last") is not None else None
            w5.close()
            return sec
        c5 = guarded("c5", fc5, log)

    # ================= oracle-anchored parity of the timed path (C1, sharded over all ranks when N > 1)
    parity_cpu = None
    if "parity" not in skip:
        parity_cpu = guarded("parity_vs_cpu_oracle", lambda: oracle_parity_c1(ctx, 32, variant), log)

    if rank == 0:
        steps = args.steps
        launches_per_step = 3 + nl1 - 1 + (1 if world > 1 else 0)  # prepare, scan(s), re-score+select(+push|+auto-merge), (merge+auto-merge)
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "serial": dict(head["serial"], note="same K steps with no overlap between consecutive steps (one stream, eager launches)"),
            "serial_graph": head.get("serial_graph"),
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.tag}: exact cosine top-{TOP_K} + auto-merge, {args.rows} x {DIM} bf16 leaf embeddings, "
                                   f"{LEVELS}-level tree, batch-1 queries, corpus row-sharded over {world} GPU(s)",
                       "rows_per_gpu": head["rows_per_gpu"], "batch": 1, "k": TOP_K, "kprime": args.kprime, "variant": args.variant,
                       "l2": "no flush: every step streams the whole shard (>= 2.5 GB) through a 126 MB L2",
                       "pipeline": "2 lanes (streams), one CUDA graph of the whole step per lane: step i+1's scan overlaps step i's "
                                   "tail (and the drain of scan i overlaps the ramp of scan i+1, so ms_per_step can sit a hair below kernel_ms)",
                       "timing": "CUDA events on the launching stream, after a device-side rendezvous of all ranks (tt_peer_barrier); max over ranks",
                       "seed": SEED},
            "roofline": head["roofline"],
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": steps * launches_per_step,
            "kernels_per_step": launches_per_step,
            "clocks": pip1["clocks"],
            "batch64": batch64,
            "wide": wide,
            "fp32_store": fp32_store,
            "hard_queries": hard,
            "c3": c3, "c4": c4, "c5": c5,
            "certificate_failures": head["certificate_failures"], "min_margin": head["min_margin"], "eps": head["eps"],
            "exchange": None if world == 1 else ("peer: fused into the select / merge kernels over symmetric memory, device-resident epochs, CUDA-graph replay"
                                                 if os.environ.get("TT_EXCHANGE", "peer") != "nccl" else "nccl all-gather"),
            "parity_vs_gpu_exact_scan": parity_ok,
            "parity_vs_cpu_oracle": parity_cpu,
            "bench_seconds": round(time.perf_counter() - t_start, 1),
        }
        print(json.dumps(line))
    if world > 1:
        ctx.dist.destroy_process_group()


# --------------------------------------------------------------------------- the CPU arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cport
    from tensor_truth_b200.synth import SynthCorpus

    cport.build()
    cport.use_all_host_threads()  # torchrun exports OMP_NUM_THREADS=1 to its workers; this arm is the all-cores baseline
    n_rows = args.rows
    full = host_can_hold(n_rows * DIM * 2) and not args.cpu_sample_rows
    sample = n_rows if full else min(args.cpu_sample_rows or 1_048_576, n_rows)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    sc = SynthCorpus(n_rows, DIM, LEVELS, SEED, device=dev)
    corpus, inv = sc.rows(0, sample)
    if full:  # the same 64 queries the GPU arm draws
        q = sc.finish_queries(sc.queries(QUERY_POOL, lookup=lambda t: corpus[t])).numpy()
    else:     # keep the targets inside the sample
        tgt = sc.query_targets(QUERY_POOL) % sample
        q = torch.zeros((QUERY_POOL, DIM))
        for j in range(QUERY_POOL):
            g = torch.Generator(device="cpu")
            g.manual_seed(SEED * 7_368_787 + 11 + j)
            q[j] = corpus[int(tgt[j])].float().cpu() + (0.3 / DIM ** 0.5) * torch.randn(DIM, generator=g)
        q = sc.finish_queries(q).numpy()
    bits = to_host_bits(corpus) if corpus.is_cuda else corpus.view(torch.int16).numpy().view(np.uint16)
    inv_h = inv.cpu().numpy()
    del corpus
    if dev == "cuda":
        torch.cuda.empty_cache()
    # calibrate so that warmup + steps stay within a few minutes
    sec, _ = cpu_arm(bits, inv_h, sc.tree, q, TOP_K, 1, 1)
    budget = 150.0
    if sec * (args.steps + args.warmup) > budget:
        sample = max(65536, int(sample * budget / (sec * (args.steps + args.warmup))) // 1024 * 1024)
        bits, inv_h = bits[:sample], inv_h[:sample]
        full = False
    sec, done = cpu_arm(bits, inv_h, sc.tree, q, TOP_K, args.steps, args.warmup, budget_s=budget)
    scale = sample / n_rows
    value = (1.0 / sec) * scale
    cores = cport.fast_threads()
    desc = (f"{done} batch-1 queries over " + (f"ALL {sample} rows" if full else f"the first {sample} of {n_rows} rows, q/s scaled by {scale:.4f} (the scan is linear in rows)")
            + f" ({sec * 1e3:.1f} ms each); oracle/c/oracle_fast.c fp32 AVX2 + OpenMP on {cores} threads, "
            f"auto-merge in oracle/automerge.py.  The reference's own stack (llama-index + chromadb) is not installable here.")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
            "warmup": args.warmup, "ms_per_step": sec * 1e3 / scale, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C2: exact cosine top-{TOP_K} + auto-merge, {n_rows} x {DIM} bf16 leaf embeddings, "
                                   f"{LEVELS}-level tree, batch-1 queries" + ("" if full else f" (CPU arm on a {sample}-row sample)"),
                       "batch": 1, "k": TOP_K, "seed": SEED, "full_corpus": bool(full)},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                             "full_corpus": bool(full), "host_cpus": os.cpu_count()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=N_ROWS)
    ap.add_argument("--kprime", type=int, default=0, help="per-CTA shortlist length (0 = by k)")
    ap.add_argument("--k", type=int, default=10, help="similarity_top_k (BASELINE: 10; C5: 200)")
    ap.add_argument("--levels", type=int, default=3, help="levels of the node tree (C5: 4)")
    ap.add_argument("--skip", default="", help="comma list of secondary sections to skip: batch64,wide,e2e,cpu,hard,fp32,c3,c4,c5,parity")
    ap.add_argument("--only-headline", action="store_true", help="skip every secondary section")
    ap.add_argument("--c3-rows-per-gpu", type=int, default=12_500_000)
    ap.add_argument("--c4-rows-per-gpu", type=int, default=6_250_000)
    ap.add_argument("--c5-rows", type=int, default=20_000_000)
    ap.add_argument("--variant", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--cpu-sample-rows", type=int, default=0,
                    help="CPU arm: 0 = the whole corpus when host memory allows (else 1,048,576 rows); N = the first N rows, q/s scaled")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="label only: 'strong' = --rows is the total (default: the same 10M rows at every N); 'weak' when the "
                         "caller scales --rows with N (C3: 12.5M rows per GPU)")
    ap.add_argument("--tag", default="C2", help="BASELINE config label written into config.workload")
    ap.add_argument("--wide-batch", type=int, default=16384, help="queries per step of the wide-batch (C4-shaped) section; 0 = skip")
    ap.add_argument("--wide-k", type=int, default=100)
    ap.add_argument("--wide-steps", type=int, default=2)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.only_headline:
        args.skip = "batch64,wide,e2e,cpu,hard,fp32,c3,c4,c5,parity"
    global TOP_K, LEVELS, METRIC
    TOP_K, LEVELS = args.k, args.levels
    if args.kprime <= 0:  # per-CTA shortlist length: deeper for larger k (clustered hits share a tile, hence a CTA)
        args.kprime = 32 if TOP_K <= 16 else 64 if TOP_K <= 32 else 128
    if (TOP_K, LEVELS, args.rows) != (10, 3, N_ROWS):
        METRIC = (f"queries/sec exact top-{TOP_K} (+auto-merge, {LEVELS}-level tree) over {args.rows} x {DIM} chunks, batch-1")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
