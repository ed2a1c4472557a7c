"""Row-sharded index: one process per GPU, each holding a contiguous block of corpus rows.

Not in the reference (single process, single device); SURVEY.md 8e.  Per query batch every rank
computes the exact top-k of its own rows, the per-rank ``(key, id)`` records are exchanged with ONE
all-gather (NCCL over NVLink on the GPU box, gloo in the CPU tests), every rank runs the same k-way
merge (key desc, id asc) and the auto-merge runs once on the merged list (tree arrays replicated).

Record layout per rank (what travels):  ``[ keys float32 B*k | pad to 8 B | ids int64 B*k ]``.
"""

from __future__ import annotations

import contextlib
import os
import threading
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def record_layout(b: int, k: int) -> Tuple[int, int, int]:
    """(record bytes, byte offset of the ids section, bytes of the keys section incl. padding)."""
    keys_bytes = (b * k * 4 + 7) // 8 * 8
    return keys_bytes + b * k * 8, keys_bytes, keys_bytes


def shard_bounds(n_rows: int, world: int, rank: int, align: int = 128) -> Tuple[int, int]:
    """Contiguous row range of ``rank``; interior cuts on multiples of ``align`` rows."""
    def cut(r):
        if r >= world:
            return n_rows
        return min(n_rows, (n_rows * r // world) // align * align)
    return cut(rank), cut(rank + 1)


class ShardedSearch:
    """Gather + merge plumbing, independent of where the local top-k comes from.

    ``local_search(q, k, keys_out, ids_out, slot)`` fills this rank's exact top-k (keys float32 [B,k], ids int64
    [B,k], global ids) into the given views of the send record; ``merge(recv, world, b, k, k_out, slot)`` returns
    ``(scores float32 [B,k_out], ids int64 [B,k_out])`` from the gathered records ``recv`` (uint8 [world, rec]).
    ``slot`` selects one of several independent buffer sets (one per CUDA stream in a pipelined caller)."""

    def __init__(self, local_search: Callable, merge: Callable, device: torch.device, group=None):
        self.local_search, self.merge, self.device, self.group = local_search, merge, device, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._bufs: dict = {}

    def buffers(self, b: int, k: int, slot: int = 0):
        w = self._bufs.get((b, k, slot))
        if w is None:
            rec, ids_off, _ = record_layout(b, k)
            send = torch.zeros(rec, dtype=torch.uint8, device=self.device)
            recv = torch.zeros((self.world, rec), dtype=torch.uint8, device=self.device)
            keys = send[: b * k * 4].view(torch.float32).view(b, k)
            ids = send[ids_off: ids_off + b * k * 8].view(torch.int64).view(b, k)
            w = self._bufs[(b, k, slot)] = (send, recv, keys, ids)
        return w

    def exchange(self, send: torch.Tensor, recv: torch.Tensor) -> None:
        if self.world > 1:
            dist.all_gather_into_tensor(recv.view(-1), send, group=self.group)
        else:
            recv[0].copy_(send)

    def search(self, q: torch.Tensor, k: int, k_out: Optional[int] = None, slot: int = 0):
        b = int(q.shape[0])
        send, recv, keys, ids = self.buffers(b, k, slot)
        self.local_search(q, k, keys, ids, slot)
        self.exchange(send, recv)
        return self.merge(recv, self.world, b, k, k_out or k, slot)


class PeerBuffers:
    """Symmetric (peer-mapped) receive regions + flags for one (batch, k) shape: what lets the selecting kernel push
    its record straight into every rank's memory and the merging kernel wait on flags (include/tt_b200.h,
    ``tt_exchange_t``).

    ``lanes`` independent lanes (one per stream of a pipelined caller), each a ring of ``SLOTS`` = 2 receive regions
    and its own device-resident epoch counter: the kernels read the epoch -- and with it the slot -- from device
    memory, so the descriptor of a lane never changes and a captured CUDA graph of the step replays as it is.  Two
    slots per lane suffice: slot e % 2 is rewritten at epoch e + 2, after this rank merged e + 1, which needed every
    peer's push of e + 1, which that peer issued (stream order) after ITS merge of e -- nobody still reads the slot.
    All ranks must issue the same sequence of pushes per lane.  One extra flag array serves ``barrier()``."""

    SLOTS = 2

    def __init__(self, b: int, k: int, device: torch.device, group=None, lanes: int = 2, _local=None):
        self.b, self.k, self.lanes, self.device = b, k, lanes, device
        if _local is not None:  # (world, rank, [buffers of all simulated ranks]): see local_group()
            self.world, self.rank, bufs = _local
        else:
            self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        rec, self.ids_off, _ = record_layout(b, k)
        self.margins_off = rec               # [keys | ids | margins float32 B]: the source's certificate margins ride along
        self.rec_bytes = rec + 4 * b         # what one source writes (and what tt_exchange_push copies)
        self.rec_stride = (self.rec_bytes + 15) // 16 * 16
        self.region = self.world * self.rec_stride                      # one slot
        self.flag_stride = (self.world + 31) // 32 * 32                 # uint32 elements per slot
        self.flags_off = (lanes * self.SLOTS * self.region + 127) // 128 * 128
        n_flag_arrays = lanes * self.SLOTS + 1                          # + the barrier's
        self.total_bytes = self.flags_off + n_flag_arrays * self.flag_stride * 4
        if _local is not None:
            self.buf = bufs[self.rank]
            self.ptrs = [int(t.data_ptr()) for t in bufs]
        else:
            import torch.distributed._symmetric_memory as symm_mem

            self.buf = symm_mem.empty(self.total_bytes, dtype=torch.uint8, device=device)
            self.buf.zero_()
            self.hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
            self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self._bind()
        torch.cuda.synchronize(device)
        if _local is None:
            dist.barrier(group)  # every rank's zero-fill has landed before anyone pushes

    @classmethod
    def local_group(cls, world: int, b: int, k: int, device: torch.device, lanes: int = 2):
        """``world`` simulated ranks inside ONE process on ONE device: plain device buffers stand in for the peer
        mappings (a rank's "peer pointer" is just the other rank's buffer).  The kernels cannot tell the difference --
        same descriptors, same stores, same flags -- which is what lets a single-GPU box exercise the exchange
        protocol (tests/test_gpu_exchange_one_device.py).  The caller must run the ranks on different streams."""
        stride = (record_layout(b, k)[0] + 4 * b + 15) // 16 * 16
        flags_off = (lanes * cls.SLOTS * world * stride + 127) // 128 * 128
        total = flags_off + (lanes * cls.SLOTS + 1) * ((world + 31) // 32 * 32) * 4
        bufs = [torch.zeros(total, dtype=torch.uint8, device=device) for _ in range(world)]
        return [cls(b, k, device, lanes=lanes, _local=(world, r, bufs)) for r in range(world)]

    def _bind(self) -> None:
        from ._lib import Exchange

        lanes, device = self.lanes, self.device
        self.epochs = torch.zeros(lanes + 1, dtype=torch.int32, device=device)   # device-resident epoch per lane (+ barrier)
        self.tickets = torch.zeros(lanes + 1, dtype=torch.int32, device=device)
        self.all_margins = torch.zeros((lanes, self.world, self.b), dtype=torch.float32, device=device)
        self._x = []
        for lane in range(lanes + 1):
            x = Exchange()
            x.world, x.rank, x.epoch = self.world, self.rank, 0
            x.rec_stride_bytes, x.ids_off_bytes = self.rec_stride, self.ids_off
            x.margins_off_bytes = self.margins_off
            barrier = lane == lanes
            for p in range(self.world):
                flags = self.ptrs[p] + self.flags_off + lane * self.SLOTS * self.flag_stride * 4
                x.peer_flags[p] = flags
                x.peer_recv[p] = flags if barrier else self.ptrs[p] + lane * self.SLOTS * self.region
            x.ticket = self.tickets.data_ptr() + 4 * lane
            x.epoch_dev = self.epochs.data_ptr() + 4 * lane
            x.n_slots = 1 if barrier else self.SLOTS
            x.slot_stride_bytes = 0 if barrier else self.region
            x.flag_slot_stride = self.flag_stride
            self._x.append(x)

    def desc(self, lane: int = 0):
        """The (constant) exchange descriptor of a lane; the library copies it into the kernel parameters at launch."""
        return self._x[lane]

    def region_ptr(self, lane: int = 0) -> int:
        """Local address of the lane's receive ring (slot 0); the kernels add the slot offset themselves."""
        return self.ptrs[self.rank] + lane * self.SLOTS * self.region

    def barrier(self, lib, stream) -> None:
        """Device-side rendezvous of all ranks, enqueued on ``stream`` (``tt_peer_barrier``): no host involved, so what
        follows it in stream order starts within an NVLink round trip of the slowest rank."""
        import ctypes as C

        from ._lib import check

        check(lib.tt_peer_barrier(C.byref(self._x[self.lanes]), stream))

    def send_record(self, lane: int = 0):
        """A local record in the layout of one source's slice of a receive region (for tt_exchange_push), one per lane:
        the uint8 buffer and its keys [B,k] / ids [B,k] / margins [B] views."""
        if not hasattr(self, "_send"):
            self._send = {}
        if lane not in self._send:
            buf = torch.zeros(self.rec_stride, dtype=torch.uint8, device=self.buf.device)
            b, k = self.b, self.k
            self._send[lane] = (buf, buf[: b * k * 4].view(torch.float32).view(b, k),
                                buf[self.ids_off: self.ids_off + b * k * 8].view(torch.int64).view(b, k),
                                buf[self.margins_off: self.margins_off + 4 * b].view(torch.float32))
        return self._send[lane]


class ShardedIndex:
    """``DeviceIndex`` shards + the fused peer exchange (or an NCCL all-gather) + the CUDA merge and auto-merge kernels."""

    LANES = 2

    def __init__(self, local_index, group=None, _local=None):
        """``_local = (world, rank, registry)``: one of several simulated ranks inside ONE process on ONE device
        (``LocalShardGroup``): the exchange buffers come from ``PeerBuffers.local_group`` instead of symmetric memory."""
        from . import _lib

        self._lib = _lib
        self.local = local_index
        self.device = local_index.device
        self._margins = None
        self._local = _local
        self._lane_locks = [threading.Lock() for _ in range(self.LANES)]
        self._lane_streams: dict = {}
        self.plumbing = ShardedSearch(self._local_search, self._merge, self.device, group)
        if _local is not None:
            self.plumbing.world, self.plumbing.rank = _local[0], _local[1]
        elif self.plumbing.world > 1:
            # every rank decides from the SAME margins whether a second round is due, so all must compare them with the same
            # eps: a measured store eps (fp32 master, index._shadow_gap) differs per shard -- take the largest
            e = torch.tensor([float(local_index.eps)], dtype=torch.float64, device=self.device)
            dist.all_reduce(e, op=dist.ReduceOp.MAX, group=group)
            local_index.eps = float(e.item())
        self._out: dict = {}
        self.group = group
        self._peers: dict = {}
        self._host_bufs: dict = {}
        self._graphs: dict = {}
        self.second_rounds = 0
        # peer pushes need symmetric memory (NVLink peer mappings); without it the exchange is an NCCL all-gather
        self.transport = "nccl" if os.environ.get("TT_EXCHANGE", "peer") == "nccl" or self.plumbing.world == 1 else "peer"

    def peers(self, b: int, k: int) -> Optional[PeerBuffers]:
        if self.transport != "peer":
            return None
        pb = self._peers.get((b, k))
        if pb is None and self._local is not None:
            world, rank, registry = self._local
            group = registry.get((b, k))
            if group is None:
                group = registry[(b, k)] = PeerBuffers.local_group(world, b, k, self.device, lanes=self.LANES)
            pb = self._peers[(b, k)] = group[rank]
        if pb is None:
            try:
                pb = self._peers[(b, k)] = PeerBuffers(b, k, self.device, self.group, lanes=self.LANES)
            except Exception as exc:  # every rank fails alike (same node, same driver): fall back together
                import warnings

                warnings.warn(f"symmetric memory unavailable ({exc}); exchanging with NCCL all-gather")
                self.transport = "nccl"
                return None
        return pb

    def _outputs(self, b, k_out, slot):
        o = self._out.get((b, k_out, slot))
        if o is None:
            o = self._out[(b, k_out, slot)] = (torch.empty((b, k_out), dtype=torch.float32, device=self.device),
                                               torch.empty((b, k_out), dtype=torch.int64, device=self.device))
        return o

    def _merge_pulled(self, pb, lane, b, k, k_out, slot, am=None, all_margins=None):
        """Merge straight out of this rank's receive ring once every source's flag has reached the lane's epoch
        (``tt_merge_topk_fused``); the same kernel copies out the margins every source pushed (``all_margins``
        float32 [world, B]) and runs the auto-merge on the merged list (``am``)."""
        import ctypes as C

        L, ptr, check = self._lib.lib(), self._lib.ptr, self._lib.check
        o = self._outputs(b, k_out, slot)
        x = pb.desc(lane)
        region = pb.region_ptr(lane)
        with self.local._on_device():
            check(L.tt_merge_topk_fused(region, region + pb.ids_off, pb.world, pb.rec_stride // 4, pb.rec_stride // 8, b, k,
                                        k_out, self.local.score_mode, ptr(o[0]), ptr(o[1]), C.byref(x),
                                        ptr(all_margins) if all_margins is not None else None,
                                        C.byref(am) if am is not None else None, self.local._stream()))
        return o

    def _local_search(self, q, k, keys_out, ids_out, slot=0):
        w = dict(self.local._buffers(int(q.shape[0]), k, slot))
        w["keys"], w["ids"] = keys_out, ids_out
        if self._margins is not None:
            w["margin"] = self._margins
        self.last = self.local.search(q, k, out=w)
        return self.last

    def _merge(self, recv, world, b, k, k_out, slot=0):
        L, ptr, check = self._lib.lib(), self._lib.ptr, self._lib.check
        rec, ids_off, _ = record_layout(b, k)
        o = self._outputs(b, k_out, slot)
        base = recv.data_ptr()
        with self.local._on_device():
            check(L.tt_merge_topk(base, base + ids_off, world, rec // 4, rec // 8, b, k, k_out,
                                  self.local.score_mode, ptr(o[0]), ptr(o[1]), self.local._stream()))
        return o

    @staticmethod
    def _run(gen):
        """Run a phased procedure (a generator that yields wherever a wait for the peers follows a push) to its end."""
        try:
            while True:
                next(gen)
        except StopIteration as stop:
            return stop.value

    def _search_phases(self, q, k, margins, slot, merged_out, ratio_thresh):
        self._margins = margins
        q = self.local._check_queries(q)
        b = int(q.shape[0])
        pb = self.peers(b, k)
        if pb is None:
            scores, ids = self.plumbing.search(q, k, slot=slot)
            if merged_out is not None:
                self.local.automerge(ids, scores, ratio_thresh, out=merged_out)
            return scores, ids
        # fused exchange: the selecting kernel pushes to every peer, the merging kernel waits on the flags
        w = dict(self.local._buffers(b, k, slot))
        if margins is not None:
            w["margin"] = margins
        self.last = self.local.search(q, k, out=w, xchg=pb.desc(slot))
        yield  # every rank's push is enqueued before anybody enqueues a wait (matters to LocalShardGroup only)
        am = self.local._am_args(ratio_thresh, merged_out) if merged_out is not None else None
        return self._merge_pulled(pb, slot, b, k, k, slot, am=am)

    def search(self, q, k, margins: Optional[torch.Tensor] = None, slot: int = 0, merged_out=None, ratio_thresh: float = 0.5):
        """Merged exact top-k, identical on every rank: ``(scores [B,k], ids [B,k])``.  ``margins`` (float32 [B])
        receives this rank's certificate margins (compare with ``self.last.eps``).  ``slot``: the lane (stream) of a
        pipelined caller, < ``LANES``.  ``merged_out`` (``MergeResult``): also auto-merge the merged list into it -- on
        the peer transport inside the merging kernel."""
        return self._run(self._search_phases(q, k, margins, slot, merged_out, ratio_thresh))

    def step_graph(self, b: int, k: int, ratio_thresh: float = 0.5, lane: int = 0, merged: bool = True):
        """The sharded step over device-resident buffers as ONE CUDA graph: prepare -> scan -> re-score + select + push
        -> flag-wait + merge + auto-merge (4 kernel nodes for batch <= 32; the exchange lives inside the last two).
        Every rank must create and replay its graphs in the same order per lane.  Peer transport only."""
        from .index import MergeResult, StepGraph, capture_on_side_stream

        key = (b, k, float(ratio_thresh), lane, merged)
        g = self._graphs.get(key)
        if g is None:
            pb = self.peers(b, k)
            if pb is None or self._local is not None:
                raise RuntimeError("ShardedIndex.step_graph needs the peer transport (symmetric memory) and one process per rank")
            dev = self.device
            q = torch.zeros((b, self.local.dim), dtype=torch.float32, device=dev)
            margins = torch.empty((b,), dtype=torch.float32, device=dev)
            mo = None
            if merged:
                mo = MergeResult(torch.empty((b, max(2 * k, 1)), dtype=torch.int64, device=dev),
                                 torch.empty((b, max(2 * k, 1)), dtype=torch.float64, device=dev),
                                 torch.empty((b,), dtype=torch.int32, device=dev))
            out = {}

            def fn():
                scores, ids = self.search(q, k, margins=margins, slot=lane, merged_out=mo, ratio_thresh=ratio_thresh)
                out["r"] = (scores, ids)
                return self.last

            with self.local._on_device():
                graph, last = capture_on_side_stream(dev, fn)
            from .index import SearchResult

            scores, ids = out["r"]
            r = SearchResult(scores, scores, ids, margins, last.eps, last.hi_only)  # merged keys are reported as scores
            g = self._graphs[key] = StepGraph(graph, q, r, mo, last.eps)
        return g

    def barrier(self, b: int = 1, k: int = 10) -> None:
        """Device-side rendezvous of all ranks on the current stream (peer transport); a host barrier otherwise."""
        if self._local is not None:
            return  # simulated ranks share one process and one device: nothing to rendezvous
        pb = self.peers(b, k)
        if pb is None:
            if self.plumbing.world > 1:
                dist.barrier(self.group)
            return
        with self.local._on_device():
            pb.barrier(self._lib.lib(), self.local._stream())

    @property
    def tree(self):
        return self.local.tree

    def close(self) -> None:
        self._graphs.clear()
        self._peers.clear()
        self._out.clear()
        self._host_bufs.clear()
        self.plumbing._bufs.clear()
        self.local.close()

    def _finish_host(self, mids, scores, b, k, ratio_thresh, merged, extra=None, lane=0):
        """Auto-merge (or copy) the merged top-k into the result record and start its D2H copy; ``extra`` = (device
        tensor, pinned host tensor) copied along.  The caller synchronises on the record's event."""
        local = self.local
        rec = local._record(b, k, merged, lane=lane)
        d = rec["d"]
        if merged:
            from .index import MergeResult

            local.automerge(mids, scores, ratio_thresh, out=MergeResult(d["ids"], d["scores"], d["lens"]))
        else:
            d["ids"].copy_(mids)
            d["scores"].copy_(scores)
        rec["host"].copy_(rec["dev"], non_blocking=True)
        if extra is not None:
            extra[1].copy_(extra[0], non_blocking=True)
        rec["event"].record()
        return rec

    @staticmethod
    def _unpack_host(rec, merged):
        h = rec["hn"]
        ids_h, scores_h = h["ids"].copy(), h["scores"].astype("float64")
        lens = h["lens"].copy() if merged else (ids_h >= 0).sum(axis=1).astype("int32")
        return ids_h, scores_h, lens

    # ------------------------------------------------------------------ host in / host out, peer transport
    def _host_push(self, pb, q, rec, b, k, lane=0):
        """First half of a host round on a lane: local scan + stage 2, which pushes this rank's record (and margins) to
        every peer; this rank's margins land in the result record."""
        w = dict(self.local._buffers(b, k, slot=("host", lane)))
        w["margin"] = rec["d"]["margin"]
        return self.local.search(q, k, out=w, xchg=pb.desc(lane))

    def _host_merge(self, pb, rec, b, k, ratio_thresh, merged, lane=0):
        """Second half: flag-wait + merge (+ auto-merge), everything landing in the result record ``rec`` -- merged lists
        and the margins EVERY rank pushed (``extra``: [world, B])."""
        import ctypes as C

        from .index import MergeResult

        local = self.local
        d = rec["d"]
        L, ptr, check = self._lib.lib(), self._lib.ptr, self._lib.check
        region = pb.region_ptr(lane)
        if merged:
            o = self._outputs(b, k, ("host", lane))
            am = local._am_args(ratio_thresh, MergeResult(d["ids"], d["scores"], d["lens"]))
        else:
            o, am = (d["scores"], d["ids"]), None
        with local._on_device():
            check(L.tt_merge_topk_fused(region, region + pb.ids_off, pb.world, pb.rec_stride // 4, pb.rec_stride // 8, b, k,
                                        k, local.score_mode, ptr(o[0]), ptr(o[1]), C.byref(pb.desc(lane)), ptr(d["extra"]),
                                        C.byref(am) if am is not None else None, local._stream()))

    def _host_round(self, pb, q, rec, b, k, ratio_thresh, merged, lane=0):
        """Both halves back to back (what the captured graph of the host path holds)."""
        r = self._host_push(pb, q, rec, b, k, lane)
        self._host_merge(pb, rec, b, k, ratio_thresh, merged, lane)
        return r

    def _host_graph(self, pb, b, k, ratio_thresh, merged, lane=0):
        """The first round of ``retrieve_host`` as ONE CUDA graph: H2D of the queries from a pinned staging buffer ->
        the four kernels of ``_host_round`` -> D2H of the result record.  Captured after ``GRAPH_AFTER`` eager calls of
        the (shape, lane) -- every rank counts alike, so all ranks switch on the same call of that lane."""
        from .index import GRAPH_AFTER, capture_on_side_stream

        key = ("host", b, k, float(ratio_thresh), merged, lane)
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = {"calls": 0, "graph": None, "dead": bool(os.environ.get("TT_NO_GRAPH"))}
        if g["graph"] is not None:
            return g
        g["calls"] += 1
        if g["dead"] or g["calls"] <= GRAPH_AFTER:
            return None
        local = self.local
        rec = local._record(b, k, merged, extra_f32=pb.world * b, lane=lane)
        q_pin = torch.zeros((b, local.dim), dtype=torch.float32).pin_memory()
        q_dev = torch.zeros((b, local.dim), dtype=torch.float32, device=self.device)
        out = {}

        def fn():
            q_dev.copy_(q_pin, non_blocking=True)
            out["r"] = self._host_round(pb, q_dev, rec, b, k, ratio_thresh, merged, lane)
            rec["host"].copy_(rec["dev"], non_blocking=True)

        # the eager run inside capture_on_side_stream is a real exchange round (with whatever q_pin holds: zeros);
        # every rank makes it, so the lane's epochs stay in lockstep.  The capture never waits for the device
        # (index.capturing), so holding the capture lock cannot stall a peer
        with local._on_device():
            graph, _ = capture_on_side_stream(self.device, fn)
        g.update(graph=graph, q_pin=q_pin, q_pin_np=q_pin.numpy(), q_dev=q_dev, result=out["r"], rec=rec)
        return g

    def _host_phases(self, pb, q_host, b, k, ratio_thresh, merged, lane=0):
        """Peer transport, ONE host synchronisation per query batch: the selecting kernel pushes this rank's record AND
        its certificate margins to every peer, the merging kernel copies all ranks' margins into the result record, so
        they come back with the answer.  Only if some rank's top-k was not proven (every rank sees that, from identical
        data) do all ranks take a second round: local repair, plain push, merge again.
        A generator: it yields wherever a wait for the peers follows a push, so that ``LocalShardGroup`` can step several
        simulated ranks in lockstep; a real rank just runs it through (``_run``)."""
        import ctypes as C

        local = self.local
        g = self._host_graph(pb, b, k, ratio_thresh, merged, lane) if self._local is None else None
        if g is not None:
            if q_host.dtype == torch.float32 and q_host.device.type == "cpu":
                g["q_pin_np"][...] = q_host.numpy()  # the common case: a plain memory copy into the pinned staging buffer
            else:
                g["q_pin"].copy_(q_host)
            q, r, rec = g["q_dev"], g["result"], g["rec"]
            with local._on_device():
                g["graph"].replay()
        else:
            q = local._check_queries(q_host.to(self.device, torch.float32, non_blocking=True))
            rec = local._record(b, k, merged, extra_f32=pb.world * b, lane=lane)
            r = self._host_push(pb, q, rec, b, k, lane)
            yield
            self._host_merge(pb, rec, b, k, ratio_thresh, merged, lane)
            rec["host"].copy_(rec["dev"], non_blocking=True)
        self.last = r
        self._lib.check(self._lib.lib().tt_stream_synchronize(local._stream()))
        self._lib.check_status(local._dev_index)
        all_h = rec["hn"]["extra"].reshape(pb.world, b)
        proven = all_h > r.eps                    # [world, B], the same on every rank
        if not proven.all():
            mine = (~proven[pb.rank]).nonzero()[0]
            if mine.size:
                with local._repair_lock:  # the ladder's workspaces are shared between lanes; nothing in it waits for a peer
                    local._repair(q, k, r, torch.from_numpy(mine).to(self.device), hi_lo_first=r.hi_only)
                    self._lib.check(self._lib.lib().tt_stream_synchronize(local._stream()))
            send, s_keys, s_ids, s_margins = pb.send_record(lane)
            s_keys.copy_(r.keys)
            s_ids.copy_(r.ids)
            s_margins.fill_(float("inf"))         # what is sent now is exact (proven before, or repaired)
            with local._on_device():
                self._lib.check(self._lib.lib().tt_exchange_push(send.data_ptr(), pb.rec_bytes // 4 * 4, C.byref(pb.desc(lane)),
                                                                 local._stream()))
            yield
            scores, mids = self._merge_pulled(pb, lane, b, k, k, ("host2", lane))
            rec = self._finish_host(mids, scores, b, k, ratio_thresh, merged, lane=lane)
            rec["event"].synchronize()
            self._lib.check_status(local._dev_index)
            self.second_rounds += 1
        return self._unpack_host(rec, merged)

    def _retrieve_host_phases(self, q_host, k, ratio_thresh, merge, lane=0):
        """``retrieve_host`` as a phased procedure (peer transport), or None when this index exchanges through NCCL."""
        merged = bool(merge and self.local.tree is not None)
        if q_host.dim() != 2 or q_host.shape[1] != self.local.dim:
            raise ValueError(f"queries must be [B, {self.local.dim}], got {tuple(q_host.shape)}")
        b = int(q_host.shape[0])
        if self.transport == "peer":
            pb0 = self.peers(b, k)
            if pb0 is not None:
                return self._host_phases(pb0, q_host, b, k, ratio_thresh, merged, lane)
        return None

    def _lane_ctx(self, lane: int):
        """Lock + stream of a host lane: lane 0 runs on the caller's current stream, the others on their own."""
        if not 0 <= lane < self.LANES:
            raise ValueError(f"lane {lane}: this index has {self.LANES} lanes")
        stack = contextlib.ExitStack()
        stack.enter_context(self._lane_locks[lane])
        if lane:
            st = self._lane_streams.get(lane)
            if st is None:
                st = self._lane_streams[lane] = torch.cuda.Stream(self.device)
                st.wait_stream(torch.cuda.current_stream(self.device))
            stack.enter_context(torch.cuda.stream(st))
        return stack

    def retrieve_host(self, q_host, k, ratio_thresh: float = 0.5, merge: bool = True, lane: int = 0):
        """Host queries in, merged (+ auto-merged) lists out (numpy), with the certificate enforced per rank.
        Same contract as ``DeviceIndex.retrieve_host``, so the retriever classes take either; every rank must call it.

        ``lane`` (< ``LANES``): concurrent callers -- one serving thread per lane -- are pipelined, each lane with its own
        stream, exchange ring, buffers and captured graph, so one caller's exchange + host work overlaps the other's
        corpus scan.  Unlike ``DeviceIndex`` the lane is the CALLER's choice: every rank must issue the same sequence of
        queries per lane, which only the caller can arrange (thread t of every rank serves lane t).  Calls on the same lane
        are serialised.  The NCCL transport has one lane: its collectives are ordered process-wide."""
        if self.transport != "peer":
            lane = 0
        with self._lane_ctx(lane):
            return self._retrieve_host_locked(q_host, k, ratio_thresh, merge, lane)

    def _retrieve_host_locked(self, q_host, k, ratio_thresh, merge, lane):
        phases = self._retrieve_host_phases(q_host, k, ratio_thresh, merge, lane)
        if phases is not None:
            return self._run(phases)
        merged = bool(merge and self.local.tree is not None)
        b = int(q_host.shape[0])
        local = self.local
        q = local._check_queries(q_host.to(self.device, torch.float32, non_blocking=True))
        hb = self._host_bufs.get(b)
        if hb is None:  # margins on the device + a pinned host mirror + the event the certificate check waits on
            hb = self._host_bufs[b] = (torch.empty((b,), dtype=torch.float32, device=self.device),
                                       torch.empty((b,), dtype=torch.float32).pin_memory(), torch.cuda.Event())
        margins, margins_h, ev = hb
        send, recv, keys, ids = self.plumbing.buffers(b, k)
        self._margins = margins
        r = self._local_search(q, k, keys, ids)
        margins_h.copy_(margins, non_blocking=True)
        ev.record()
        ev.synchronize()
        self._lib.check_status(local._dev_index)
        bad = (~(margins_h > r.eps)).nonzero().flatten()
        if bad.numel():  # rank-local repair (writes into the send record); the exchange below is reached by every rank
            local._repair(q, k, r, bad.to(self.device), hi_lo_first=r.hi_only)
        # NCCL transport (or symmetric memory unavailable): certificate first (one sync), then all-gather + merge
        self.plumbing.exchange(send, recv)
        scores, mids = self._merge(recv, self.plumbing.world, b, k, k)
        rec = self._finish_host(mids, scores, b, k, ratio_thresh, merged)
        rec["event"].synchronize()
        return self._unpack_host(rec, merged)


class LocalShardGroup:
    """Several row shards of one corpus on ONE GPU in ONE process, each behind its own ``ShardedIndex`` and its own
    stream -- the multi-GPU protocol (push into every rank's receive ring, flags, merge, margins of every rank, second
    round after a repair) with plain device buffers standing in for the peer mappings.  What it is for: exercising
    ``ShardedIndex`` itself on a single-GPU box (tests/test_gpu_exchange_one_device.py); the kernels and descriptors
    are exactly those of the real thing.

    One thread drives all ranks, so the phased procedures of the ranks are stepped in LOCKSTEP: every rank enqueues
    its push before any rank enqueues the wait that follows -- a kernel never waits for one enqueued after it (two
    streams of one process may share a hardware queue)."""

    def __init__(self, shards):
        registry: dict = {}
        world = len(shards)
        eps = max(float(s.eps) for s in shards)
        for s in shards:
            s.eps = eps  # one eps for all ranks (see ShardedIndex.__init__)
        self.ranks = [ShardedIndex(s, _local=(world, r, registry)) for r, s in enumerate(shards)]
        self.streams = [torch.cuda.Stream(s.device) for s in shards]

    def _lockstep(self, gens):
        out = [None] * len(gens)
        live = list(range(len(gens)))
        while live:
            for r in list(live):
                with torch.cuda.stream(self.streams[r]):
                    try:
                        next(gens[r])
                    except StopIteration as stop:
                        out[r] = stop.value
                        live.remove(r)
        return out

    def search(self, q, k, merged_out=None, ratio_thresh: float = 0.5):
        """Every rank's ``ShardedIndex.search``; returns the per-rank ``(scores, ids)`` (all equal)."""
        cur = torch.cuda.current_stream(self.ranks[0].device)
        for st in self.streams:
            st.wait_stream(cur)
        outs = merged_out if merged_out is not None else [None] * len(self.ranks)
        res = self._lockstep([rk._search_phases(q, k, None, 0, outs[r], ratio_thresh) for r, rk in enumerate(self.ranks)])
        for st in self.streams:
            cur.wait_stream(st)
        return res

    def retrieve_host(self, q_host, k, ratio_thresh: float = 0.5, merge: bool = True):
        """Every rank's ``ShardedIndex.retrieve_host``; returns the per-rank ``(ids, scores, lens)`` (all equal)."""
        return self._lockstep([rk._retrieve_host_phases(q_host, k, ratio_thresh, merge) for rk in self.ranks])
