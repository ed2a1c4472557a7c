"""Row-sharded index: one process per GPU, each holding a contiguous block of corpus rows.

Not in the reference (single process, single device); SURVEY.md 8e.  Per query batch every rank
computes the exact top-k of its own rows, the per-rank ``(key, id)`` records are exchanged with ONE
all-gather (NCCL over NVLink on the GPU box, gloo in the CPU tests), every rank runs the same k-way
merge (key desc, id asc) and the auto-merge runs once on the merged list (tree arrays replicated).

Record layout per rank (what travels):  ``[ keys float32 B*k | pad to 8 B | ids int64 B*k ]``.
"""

from __future__ import annotations

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

    ``local_search(q, k, keys_out, ids_out)`` fills this rank's exact top-k (keys float32 [B,k], ids int64
    [B,k], global ids) into the given views of the send record; ``merge(recv, world, b, k, k_out)`` returns
    ``(scores float32 [B,k_out], ids int64 [B,k_out])`` from the gathered records ``recv`` (uint8 [world, rec])."""

    def __init__(self, local_search: Callable, merge: Callable, device: torch.device, group=None):
        self.local_search, self.merge, self.device, self.group = local_search, merge, device, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._bufs: dict = {}

    def buffers(self, b: int, k: int):
        w = self._bufs.get((b, k))
        if w is None:
            rec, ids_off, _ = record_layout(b, k)
            send = torch.zeros(rec, dtype=torch.uint8, device=self.device)
            recv = torch.zeros((self.world, rec), dtype=torch.uint8, device=self.device)
            keys = send[: b * k * 4].view(torch.float32).view(b, k)
            ids = send[ids_off: ids_off + b * k * 8].view(torch.int64).view(b, k)
            w = self._bufs[(b, k)] = (send, recv, keys, ids)
        return w

    def search(self, q: torch.Tensor, k: int, k_out: Optional[int] = None):
        b = int(q.shape[0])
        send, recv, keys, ids = self.buffers(b, k)
        self.local_search(q, k, keys, ids)
        if self.world > 1:
            dist.all_gather_into_tensor(recv.view(-1), send, group=self.group)
        else:
            recv[0].copy_(send)
        return self.merge(recv, self.world, b, k, k_out or k)


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

    def _local_search(self, q, k, keys_out, ids_out):
        w = dict(self.local._buffers(int(q.shape[0]), k))
        w["keys"], w["ids"] = keys_out, ids_out
        if self._margins is not None:
            w["margin"] = self._margins
        self.local.search(q, k, out=w)

    def _merge(self, recv, world, b, k, k_out):
        L, ptr, check = self._lib.lib(), self._lib.ptr, self._lib.check
        rec, ids_off, _ = record_layout(b, k)
        o = self._out.get((b, k_out))
        if o is None:
            o = self._out[(b, k_out)] = (torch.empty((b, k_out), dtype=torch.float32, device=self.device),
                                         torch.empty((b, k_out), dtype=torch.int64, device=self.device))
        import ctypes as C

        base = recv.data_ptr()
        with torch.cuda.device(self.device):
            check(L.tt_merge_topk(C.c_void_p(base), C.c_void_p(base + ids_off), world, rec // 4, rec // 8, b, k, k_out,
                                  self.local.score_mode, ptr(o[0]), ptr(o[1]), torch.cuda.current_stream().cuda_stream))
        return o

    def search(self, q, k, margins: Optional[torch.Tensor] = None):
        """Merged exact top-k, identical on every rank: ``(scores [B,k], ids [B,k])``.  ``margins`` (float32 [B])
        receives this rank's certificate margins."""
        self._margins = margins
        q = self.local._check_queries(q)
        return self.plumbing.search(q, k)

    def retrieve_host(self, q_host, k, ratio_thresh: float = 0.5):
        """Host queries in, merged + auto-merged lists out (numpy), with the certificate enforced per rank."""
        q = q_host.to(self.device, torch.float32, non_blocking=True)
        b = int(q.shape[0])
        margins = torch.empty((b,), dtype=torch.float32, device=self.device)
        send, recv, keys, ids = self.plumbing.buffers(b, k)
        self._margins = margins
        self._local_search(q, k, keys, ids)
        bad = torch.nonzero(~(margins > self.local.eps)).flatten()
        if bad.numel():  # rank-local repair; the collective below is reached by every rank either way
            self.local.fallbacks += int(bad.numel())
            ex = self.local.search_exact(q.index_select(0, bad), k)
            keys.index_copy_(0, bad, ex.keys)
            ids.index_copy_(0, bad, ex.ids)
        if self.plumbing.world > 1:
            dist.all_gather_into_tensor(recv.view(-1), send, group=self.plumbing.group)
        else:
            recv[0].copy_(send)
        scores, mids = self._merge(recv, self.plumbing.world, b, k, k)
        if self.local.tree is not None:
            m = self.local.automerge(mids, scores, ratio_thresh)
            return m.ids.cpu().numpy(), m.scores.cpu().numpy(), m.lens.cpu().numpy()
        ids_h = mids.cpu()
        return ids_h.numpy(), scores.double().cpu().numpy(), (ids_h >= 0).sum(dim=1).to(torch.int32).numpy()
