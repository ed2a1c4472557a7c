"""Row-sharded index: one process per GPU, each holding a contiguous block of corpus rows.

Not in the reference (single process, single device); SURVEY.md 8e.  Per query batch every rank
computes the exact top-k of its own rows, the per-rank ``(key, id)`` records are exchanged with ONE
all-gather (NCCL over NVLink on the GPU box, gloo in the CPU tests), every rank runs the same k-way
merge (key desc, id asc) and the auto-merge runs once on the merged list (tree arrays replicated).

Record layout per rank (what travels):  ``[ keys float32 B*k | pad to 8 B | ids int64 B*k ]``.
"""

from __future__ import annotations

import os
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
    ``tt_exchange_t``).  A ring of ``DEPTH`` slots: with steps alternating between two streams, slot e % 4 is only
    rewritten (epoch e+4) after this rank merged epoch e+2, which needed every peer's push of e+2, which on that peer
    is stream-ordered after ITS merge of epoch e -- so nobody is still reading the slot."""

    DEPTH = 4

    def __init__(self, b: int, k: int, device: torch.device, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.b, self.k = b, k
        rec, self.ids_off, _ = record_layout(b, k)
        self.margins_off = rec               # [keys | ids | margins float32 B]: the source's certificate margins ride along
        self.rec_bytes = rec + 4 * b         # what one source writes (and what tt_exchange_push copies)
        self.rec_stride = (self.rec_bytes + 15) // 16 * 16
        self.region = self.world * self.rec_stride
        self.flags_off = (self.DEPTH * self.region + 127) // 128 * 128
        total = self.flags_off + (self.DEPTH * self.world * 4 + 127) // 128 * 128
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        self.tickets = torch.zeros(self.DEPTH, dtype=torch.int32, device=device)
        self.epoch = 0
        from ._lib import Exchange

        self._x = []
        for slot in range(self.DEPTH):  # one descriptor per slot; only the epoch changes from step to step
            x = Exchange()
            x.world, x.rank = self.world, self.rank
            x.rec_stride_bytes, x.ids_off_bytes = self.rec_stride, self.ids_off
            x.margins_off_bytes = self.margins_off
            for p in range(self.world):
                x.peer_recv[p] = self.ptrs[p] + slot * self.region
                x.peer_flags[p] = self.ptrs[p] + self.flags_off + slot * self.world * 4
            x.ticket = self.tickets.data_ptr() + 4 * slot
            self._x.append(x)
        torch.cuda.synchronize(device)
        dist.barrier(group)  # every rank's zero-fill has landed before anyone pushes

    def next(self):
        """The exchange descriptor of the next step (all ranks call this in lockstep) and the local addresses the merge reads."""
        self.epoch += 1
        slot = (self.epoch - 1) % self.DEPTH
        x = self._x[slot]  # the library copies it into the kernel parameters at launch, so reuse is safe
        x.epoch = self.epoch
        self.last_slot = slot
        return x, self.ptrs[self.rank] + slot * self.region

    def margins_view(self, slot: int) -> torch.Tensor:
        """float32 [world, B] view of the margins every source pushed into this rank's receive region of ``slot``."""
        base = slot * self.region
        recs = self.buf[base: base + self.world * self.rec_stride].view(self.world, self.rec_stride)
        return recs[:, self.margins_off: self.margins_off + 4 * self.b].view(torch.float32)

    def send_record(self):
        """A local record in the layout of one source's slice of a receive region (for tt_exchange_push): the uint8
        buffer and its keys [B,k] / ids [B,k] / margins [B] views."""
        if not hasattr(self, "_send"):
            buf = torch.zeros(self.rec_stride, dtype=torch.uint8, device=self.buf.device)
            b, k = self.b, self.k
            self._send = (buf, buf[: b * k * 4].view(torch.float32).view(b, k),
                          buf[self.ids_off: self.ids_off + b * k * 8].view(torch.int64).view(b, k),
                          buf[self.margins_off: self.margins_off + 4 * b].view(torch.float32))
        return self._send


class ShardedIndex:
    """``DeviceIndex`` shards + NCCL all-gather + the CUDA merge and auto-merge kernels."""

    def __init__(self, local_index, group=None):
        from . import _lib

        self._lib = _lib
        self.local = local_index
        self.device = local_index.device
        self._margins = None
        self.plumbing = ShardedSearch(self._local_search, self._merge, self.device, group)
        self._out: dict = {}
        self.group = group
        self._peers: dict = {}
        self._host_bufs: dict = {}
        # peer pushes need symmetric memory (NVLink peer mappings); without it the exchange is an NCCL all-gather
        self.transport = "nccl" if os.environ.get("TT_EXCHANGE", "peer") == "nccl" or self.plumbing.world == 1 else "peer"

    def peers(self, b: int, k: int) -> Optional[PeerBuffers]:
        if self.transport != "peer":
            return None
        pb = self._peers.get((b, k))
        if pb is None:
            try:
                pb = self._peers[(b, k)] = PeerBuffers(b, k, self.device, self.group)
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

    def _merge_pulled(self, x, region_ptr, b, k, k_out, slot):
        """Merge straight out of this rank's receive region once every source's flag has reached the epoch."""
        import ctypes as C

        L, ptr, check = self._lib.lib(), self._lib.ptr, self._lib.check
        o = self._outputs(b, k_out, slot)
        with self.local._on_device():
            check(L.tt_merge_topk_pulled(region_ptr, region_ptr + x.ids_off_bytes, x.world, x.rec_stride_bytes // 4,
                                         x.rec_stride_bytes // 8, b, k, k_out, self.local.score_mode, ptr(o[0]), ptr(o[1]),
                                         C.byref(x), self.local._stream()))
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

    def search(self, q, k, margins: Optional[torch.Tensor] = None, slot: int = 0):
        """Merged exact top-k, identical on every rank: ``(scores [B,k], ids [B,k])``.  ``margins`` (float32 [B])
        receives this rank's certificate margins (compare with ``self.last.eps``)."""
        self._margins = margins
        q = self.local._check_queries(q)
        b = int(q.shape[0])
        pb = self.peers(b, k)
        if pb is None:
            return self.plumbing.search(q, k, slot=slot)
        # fused exchange: the selecting kernel pushes to every peer, the merging kernel waits on the flags
        x, region = pb.next()
        w = dict(self.local._buffers(b, k, slot))
        if margins is not None:
            w["margin"] = margins
        self.last = self.local.search(q, k, out=w, xchg=x)
        return self._merge_pulled(x, region, b, k, k, slot)

    @property
    def tree(self):
        return self.local.tree

    def close(self) -> None:
        self._peers.clear()
        self._out.clear()
        self._host_bufs.clear()
        self.plumbing._bufs.clear()
        self.local.close()

    def _finish_host(self, mids, scores, b, k, ratio_thresh, merged, extra=None):
        """Auto-merge (or copy) the merged top-k into the result record and start its D2H copy; ``extra`` = (device
        tensor, pinned host tensor) copied along.  The caller synchronises on the record's event."""
        local = self.local
        rec = local._record(b, k, merged)
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
        h = rec["h"]
        ids_h, scores_h = h["ids"].numpy().copy(), h["scores"].numpy().astype("float64")
        lens = h["lens"].numpy().copy() if merged else (ids_h >= 0).sum(axis=1).astype("int32")
        return ids_h, scores_h, lens

    def _retrieve_host_peer(self, pb, q, b, k, ratio_thresh, merged):
        """Peer transport, ONE host synchronisation per query batch: the selecting kernel pushes this rank's record AND
        its certificate margins to every peer, so after the merge every rank finds all ranks' margins in its own
        receive region and they come back with the result.  Only if some rank's top-k was not proven (every rank sees
        that, from identical data) do all ranks take a second round: local repair, plain push, merge again."""
        import ctypes as C

        local = self.local
        hb = self._host_bufs.get(("all", b))
        if hb is None:
            hb = self._host_bufs[("all", b)] = (torch.empty((b,), dtype=torch.float32, device=self.device),
                                                torch.empty((pb.world, b), dtype=torch.float32).pin_memory())
        margins, all_h = hb
        x, region = pb.next()
        slot = pb.last_slot
        w = dict(local._buffers(b, k))
        w["margin"] = margins
        r = self.last = local.search(q, k, out=w, xchg=x)
        scores, mids = self._merge_pulled(x, region, b, k, k, 0)
        rec = self._finish_host(mids, scores, b, k, ratio_thresh, merged, extra=(pb.margins_view(slot), all_h))
        rec["event"].synchronize()
        proven = all_h.numpy() > r.eps           # [world, B], the same on every rank
        if not proven.all():
            mine = (~proven[pb.rank]).nonzero()[0]
            if mine.size:
                local._repair(q, k, r, torch.from_numpy(mine).to(self.device), hi_lo_first=r.hi_only)
            send, s_keys, s_ids, s_margins = pb.send_record()
            s_keys.copy_(r.keys)
            s_ids.copy_(r.ids)
            s_margins.fill_(float("inf"))         # what is sent now is exact (proven before, or repaired)
            x2, region2 = pb.next()
            with local._on_device():
                self._lib.check(self._lib.lib().tt_exchange_push(send.data_ptr(), pb.rec_bytes // 4 * 4, C.byref(x2),
                                                                 local._stream()))
            scores, mids = self._merge_pulled(x2, region2, b, k, k, 0)
            rec = self._finish_host(mids, scores, b, k, ratio_thresh, merged)
            rec["event"].synchronize()
            self.second_rounds = getattr(self, "second_rounds", 0) + 1
        return self._unpack_host(rec, merged)

    def retrieve_host(self, q_host, k, ratio_thresh: float = 0.5, merge: bool = True):
        """Host queries in, merged (+ auto-merged) lists out (numpy), with the certificate enforced per rank.
        Same contract as ``DeviceIndex.retrieve_host``, so the retriever classes take either; every rank must call it."""
        merged = bool(merge and self.local.tree is not None)
        if self.transport == "peer":
            q0 = self.local._check_queries(q_host.to(self.device, torch.float32, non_blocking=True))
            pb0 = self.peers(int(q0.shape[0]), k)
            if pb0 is not None:
                return self._retrieve_host_peer(pb0, q0, int(q0.shape[0]), k, ratio_thresh, merged)
        local = self.local
        q = local._check_queries(q_host.to(self.device, torch.float32, non_blocking=True))
        b = int(q.shape[0])
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
        bad = (~(margins_h > r.eps)).nonzero().flatten()
        if bad.numel():  # rank-local repair (writes into the send record); the exchange below is reached by every rank
            local._repair(q, k, r, bad.to(self.device), hi_lo_first=r.hi_only)
        # NCCL transport (or symmetric memory unavailable): certificate first (one sync), then all-gather + merge
        self.plumbing.exchange(send, recv)
        scores, mids = self._merge(recv, self.plumbing.world, b, k, k)
        rec = self._finish_host(mids, scores, b, k, ratio_thresh, merged)
        rec["event"].synchronize()
        return self._unpack_host(rec, merged)
