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

    def __init__(self, index: int, period_s: float = 0.02):
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


def make_queries(sc, corpus, lo, hi, n_q, world):
    import torch.distributed as dist

    def lookup(t):
        return corpus[t - lo] if lo <= t < hi else None

    q = sc.queries(n_q, lookup=lookup)
    if world > 1:
        qd = q.cuda()
        dist.all_reduce(qd)  # every target row is owned by exactly one rank; the other ranks contribute zeros
        q = qd.cpu()
    return sc.finish_queries(q)


def cpu_arm(bits, inv_norm, tree, queries, k, n_steps, n_warm, budget_s=25.0):
    """The CPU restatement in its fast mode on every host thread (oracle/c/oracle_fast.c + oracle/automerge.py):
    one step = one batch-1 query over the sample rows + auto-merge.  Returns (seconds per step, steps run)."""
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


# --------------------------------------------------------------------------- this repo's arm
def run_b200(args):
    import torch.distributed as dist

    from tensor_truth_b200 import _lib
    from tensor_truth_b200.index import DeviceIndex
    from tensor_truth_b200.retriever import B200AutoMergingRetriever, B200VectorIndexRetriever
    from tensor_truth_b200.schema import QueryBundle
    from tensor_truth_b200.sharded import ShardedIndex

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    device = torch.device(f"cuda:{local_rank}")
    if world > 1:
        # NCCL writes its version banner (any NCCL_DEBUG level >= VERSION, which this image sets) and its logs to stdout;
        # rank 0 must print ONE line there
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "NONE"
        dist.init_process_group("nccl", device_id=device)
    _lib.lib()  # fail loudly if the CUDA library is missing

    n_rows = args.rows
    sc, corpus, inv, lo, hi = build_shard(n_rows, rank, world, device)
    queries = make_queries(sc, corpus, lo, hi, QUERY_POOL, world).to(device)
    variant = {"auto": _lib.SCAN_AUTO, "simt": _lib.SCAN_SIMT, "tcgen05": _lib.SCAN_TCGEN05}[args.variant]
    idx = DeviceIndex(corpus, sc.tree, inv_norm=inv, id_base=lo, device=device, kprime=args.kprime, variant=variant)
    sharded = ShardedIndex(idx) if world > 1 else None
    peak, peak_src = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from tensor_truth_b200.index import MergeResult

    def timed(batch: int, steps: int, warm: int, sample_clocks: bool, depth: int, k: int = 0, pool=None):
        """K steps of the device pipeline.  depth = 1: strictly serial, with CUDA events around every stage-1
        launch (the roofline numbers).  depth = 2: steps alternate between two streams, so step i+1's scan
        overlaps step i's re-score / select / (all-gather, merge) / auto-merge -- the throughput configuration."""
        margins = torch.full((steps + warm, batch), float("inf"), dtype=torch.float32, device=device)
        k = k or TOP_K
        pool = queries if pool is None else pool
        n_pool = max(1, int(pool.shape[0]) // batch)
        streams = [torch.cuda.Stream(device) for _ in range(depth)]
        bufs = [idx._buffers(batch, k, slot=s) for s in range(depth)]
        mouts = [MergeResult(torch.empty((batch, 2 * k), dtype=torch.int64, device=device),
                             torch.empty((batch, 2 * k), dtype=torch.float64, device=device),
                             torch.empty((batch,), dtype=torch.int32, device=device)) for _ in range(depth)]
        eps = [idx.eps]

        def one(i):
            s = i % depth
            with torch.cuda.stream(streams[s]):
                q = pool[(i % n_pool) * batch:(i % n_pool) * batch + batch]
                if sharded is None:
                    w = dict(bufs[s])
                    w["margin"] = margins[i]
                    r = idx.search(q, k, out=w)
                    eps[0] = r.eps
                    return idx.automerge(r.ids, r.scores, out=mouts[s])
                scores, ids = sharded.search(q, k, margins=margins[i], slot=s)
                eps[0] = sharded.last.eps
                return idx.automerge(ids, scores, out=mouts[s])

        cur = torch.cuda.current_stream()
        for i in range(warm):
            one(i)
        for st in streams:
            cur.wait_stream(st)
        barrier()
        idx.scan_events = [] if depth == 1 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(physical_index(local_rank)) if sample_clocks else None
        if sampler:
            sampler.__enter__()
        e0.record()
        for st in streams:
            st.wait_stream(cur)
        for i in range(steps):
            last = one(warm + i)
        for st in streams:
            cur.wait_stream(st)
        e1.record()
        barrier()
        if sampler:
            sampler.__exit__()
        ms = max_over_ranks(e0.elapsed_time(e1))
        ev = idx.scan_events
        idx.scan_events = None
        scan_ms, n_scan_launches = None, None
        if ev:
            scan_ms = max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in ev])))
            n_scan_launches = len(ev) // steps
        mt = margins[warm:]
        bad = int((~(mt > eps[0])).sum().item())
        return {"ms": ms, "scan_ms": scan_ms, "scan_calls_per_step": n_scan_launches, "bad": bad, "last": last,
                "clocks": sampler.summary() if sampler else None, "min_margin": float(mt.min().item()), "eps": eps[0]}

    # ---- headline: batch-1.  Serial loop first (per-kernel events -> roofline), then the pipelined loop (-> value).
    ser1 = timed(1, args.steps, args.warmup, False, depth=1)
    pip1 = timed(1, args.steps, args.warmup, True, depth=2)
    ms1, scan1_ms, nl1, bad1, clocks = pip1["ms"], ser1["scan_ms"], ser1["scan_calls_per_step"], ser1["bad"] + pip1["bad"], pip1["clocks"]
    value = args.steps * 1 / (ms1 / 1e3)
    local_bytes = float(hi - lo) * DIM * 2
    achieved = local_bytes / (scan1_ms / 1e3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "scan_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("rows") == hi - lo and tj.get("batch") == 1:
            traffic = tj.get("dram_bytes_per_launch")
    except Exception:
        pass

    # ---- batch-64 (same corpus, one pass of 64 hi-only queries; tensor work rises, HBM bytes per pass do not)
    skip = set(x for x in args.skip.split(",") if x)
    batch64 = None
    if "batch64" not in skip:
        steps64 = max(3, min(args.steps, 40))
        ser64 = timed(64, steps64, 3, False, depth=1)
        pip64 = timed(64, steps64, 3, False, depth=2)
        batch64 = {"value": steps64 * 64 / (pip64["ms"] / 1e3), "unit": UNIT, "ms_per_step": pip64["ms"] / steps64,
                   "steps": steps64, "serial_value": steps64 * 64 / (ser64["ms"] / 1e3), "scan_ms_per_step": ser64["scan_ms"],
                   "hbm_frac": float(hi - lo) * DIM * 2 / (ser64["scan_ms"] / 1e3) / 1e9 / peak,
                   "certificate_failures": ser64["bad"] + pip64["bad"], "min_margin": pip64["min_margin"],
                   "eps": pip64["eps"], "mode": "hi-only bf16 queries, 64 per corpus pass",
                   "kernel": "scan_gemm_kernel<64,2>" if idx._use_gemm(64) else "scan_tc2_kernel<64,2>"}

    # ---- wide batch (BASELINE configs[3] shape: 16k concurrent queries, top-100): the tensor-bound regime, served by
    #      the GEMM-shaped stage 1 (scan_gemm.cu).  Roofline: dense bf16 tensor throughput.
    wide = None
    if args.wide_batch > 0 and "wide" not in skip:
        try:
            bw, kw = args.wide_batch, args.wide_k
            qw = make_queries(sc, corpus, lo, hi, bw, world).to(device)
            tpeak, tpeak_src = measured_tensor_peak()
            serw = timed(bw, args.wide_steps, 1, False, depth=1, k=kw, pool=qw)
            flops = 2.0 * bw * float(hi - lo) * DIM
            tfl = flops / (serw["scan_ms"] / 1e3) / 1e12
            # spot check: the first 4 queries against the exact fp64 scan of the local shard
            rw = idx.search(qw, kw, out=dict(idx._buffers(bw, kw, slot=0)))
            exw = idx.search_exact(qw[:4], kw)
            wide_ok = bool(torch.equal(rw.ids[:4], exw.ids) and torch.equal(rw.scores[:4], exw.scores))
            wide = {"value": args.wide_steps * bw / (serw["ms"] / 1e3), "unit": UNIT, "batch": bw, "k": kw,
                    "ms_per_step": serw["ms"] / args.wide_steps, "steps": args.wide_steps,
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
            torch.cuda.empty_cache()
        except Exception as exc:  # never lose the headline line to the secondary section
            wide = {"error": f"{type(exc).__name__}: {exc}"}

    # ---- parity spot check inside the bench: the timed path vs the on-GPU exact fp64 scan of the same shard(s)
    qs = queries[:4]
    if sharded is None:
        r = idx.search(qs, TOP_K)
        got_ids, got_sc = r.ids.clone(), r.scores.clone()
        ex = idx.search_exact(qs, TOP_K)
        parity_ok = bool(torch.equal(got_ids, ex.ids) and torch.equal(got_sc, ex.scores))
    else:
        scores, ids = sharded.search(qs, TOP_K)
        got_ids, got_sc = ids.clone(), scores.clone()
        ex = idx.search_exact(qs, TOP_K)
        sharded.plumbing.local_search = lambda q, k, ko, io, slot=0: (ko.copy_(ex.keys), io.copy_(ex.ids))
        s2, i2 = sharded.plumbing.search(qs, TOP_K)
        parity_ok = bool(torch.equal(got_ids, i2) and torch.equal(got_sc, s2))
        sharded.plumbing.local_search = sharded._local_search
    torch.cuda.synchronize()
    last = pip1["last"]

    # ---- end to end through the public retriever API: host query in, NodeWithScore list out
    e2e_steps = max(5, min(args.steps, 100))
    q_host = queries.cpu()
    q_lists = [row.tolist() for row in q_host]  # host input as an embed model hands it over: a Python list of floats
    base_r = B200VectorIndexRetriever(idx if sharded is None else sharded, similarity_top_k=TOP_K)
    am = B200AutoMergingRetriever(base_r, None)  # over a ShardedIndex every rank makes the same call (SPMD)
    call = lambda i: am.retrieve(QueryBundle(query_str=f"q{i}", embedding=q_lists[i % QUERY_POOL]))  # noqa: E731
    e2e_warm = 8  # past DeviceIndex's GRAPH_AFTER: the one-off CUDA-graph capture of the pipeline belongs to the warm-up
    for i in range(e2e_warm):
        out = call(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        out = call(e2e_warm + i)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    n_out = len(out)
    e2e = {"value": e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": DIM * 4,
           "d2h_bytes_per_step": idx._record(1, TOP_K, True)["bytes"], "steps": e2e_steps, "warmup": e2e_warm,
           "api": "B200AutoMergingRetriever.retrieve(QueryBundle)" + ("" if sharded is None else " over ShardedIndex, every rank"),
           "fallbacks": idx.fallbacks, "nodes_returned_last": n_out}

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same corpus bytes
    cpu = None
    if world == 1 and not args.no_cpu and "cpu" not in skip:
        sample = min(args.cpu_sample_rows, hi - lo)
        bits = corpus[:sample].view(torch.int16).cpu().numpy().view(np.uint16)
        inv_h = inv[:sample].cpu().numpy()
        from oracle import cport

        cport.build()
        cport.use_all_host_threads()
        sec, done = cpu_arm(bits, inv_h, sc.tree, q_host.numpy(), TOP_K, 40, 2)
        cpu = {"value": (1.0 / sec) * (sample / n_rows), "unit": UNIT, "cores": cport.fast_threads(), "kind": "port",
               "sample": f"{done} batch-1 queries over the first {sample} of {n_rows} rows ({sec * 1e3:.1f} ms each), "
                         f"q/s scaled by {sample}/{n_rows} (the scan is linear in rows); oracle/c/oracle_fast.c "
                         f"fp32 AVX2 + OpenMP, auto-merge in oracle/automerge.py",
               "host_cpus": os.cpu_count()}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms1 / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "serial": {"value": args.steps / (ser1["ms"] / 1e3), "ms_per_step": ser1["ms"] / args.steps,
                       "note": "same K steps with no overlap between consecutive steps (one stream)"},
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.tag}: exact cosine top-{TOP_K} + auto-merge, {n_rows} x {DIM} bf16 leaf embeddings, "
                                   f"{LEVELS}-level tree, batch-1 queries, corpus row-sharded over {world} GPU(s)",
                       "rows_per_gpu": hi - lo, "batch": 1, "k": TOP_K, "kprime": args.kprime, "variant": args.variant,
                       "l2": "no flush: every step streams the whole shard (>= 2.5 GB) through a 126 MB L2",
                       "pipeline": "2 streams: step i+1's scan overlaps step i's re-score/select/merge/auto-merge",
                       "seed": SEED},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "scan_tc_kernel" if args.variant != "simt" else "scan_simt_kernel",
                         "bytes_per_launch": local_bytes, "kernel_ms": scan1_ms, "peak_source": peak_src,
                         "step_share": scan1_ms * nl1 / (ser1["ms"] / args.steps),
                         "measured_in": "serial loop, CUDA events around each stage-1 launch"},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": args.steps * (4 + nl1 + (1 if world > 1 else 0)),
            "clocks": clocks,
            "batch64": batch64,
            "wide": wide,
            "certificate_failures": bad1, "min_margin": pip1["min_margin"], "eps": pip1["eps"],
            "exchange": (sharded.transport + (" (fused into the select / merge kernels over peer memory)" if sharded.transport == "peer" else " all-gather")) if sharded is not None else None,
            "parity_vs_gpu_exact_scan": parity_ok,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


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
    sample = min(args.cpu_sample_rows, n_rows)
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    sc = SynthCorpus(n_rows, DIM, LEVELS, SEED, device=dev)
    corpus, inv = sc.rows(0, sample)
    tgt = sc.query_targets(QUERY_POOL) % sample  # keep the targets inside the sample
    q = torch.zeros((QUERY_POOL, DIM))
    for j in range(QUERY_POOL):
        g = torch.Generator(device="cpu")
        g.manual_seed(SEED * 7_368_787 + 11 + j)
        q[j] = corpus[int(tgt[j])].float().cpu() + (0.3 / DIM ** 0.5) * torch.randn(DIM, generator=g)
    q = sc.finish_queries(q).numpy()
    bits = corpus.view(torch.int16).cpu().numpy().view(np.uint16)
    inv_h = inv.cpu().numpy()
    del corpus
    # calibrate so that warmup + steps stay within a few minutes
    sec, _ = cpu_arm(bits, inv_h, sc.tree, q, TOP_K, 1, 1)
    budget = 150.0
    if sec * (args.steps + args.warmup) > budget:
        sample = max(65536, int(sample * budget / (sec * (args.steps + args.warmup))) // 1024 * 1024)
        bits, inv_h = bits[:sample], inv_h[:sample]
    sec, done = cpu_arm(bits, inv_h, sc.tree, q, TOP_K, args.steps, args.warmup, budget_s=budget)
    value = (1.0 / sec) * (sample / n_rows)
    cores = cport.fast_threads()
    desc = (f"{done} batch-1 queries over the first {sample} of {n_rows} rows ({sec * 1e3:.1f} ms each), q/s scaled by "
            f"{sample}/{n_rows} (the scan is linear in rows); oracle/c/oracle_fast.c fp32 AVX2 + OpenMP on {cores} threads, "
            f"auto-merge in oracle/automerge.py.  The reference's own stack (llama-index + chromadb) is not installable here.")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
            "warmup": args.warmup, "ms_per_step": sec * 1e3 * (n_rows / sample), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C2: exact cosine top-{TOP_K} + auto-merge, {n_rows} x {DIM} bf16 leaf embeddings, "
                                   f"{LEVELS}-level tree, batch-1 queries (CPU arm on a {sample}-row sample)",
                       "batch": 1, "k": TOP_K, "seed": SEED},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
                             "host_cpus": os.cpu_count()},
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
    ap.add_argument("--skip", default="", help="comma list of secondary sections to skip: batch64,wide,e2e,cpu")
    ap.add_argument("--variant", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--cpu-sample-rows", type=int, default=1_048_576)
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
