"""The row-sharded exchange protocol on ONE GPU: two simulated ranks in one process (two shards of the corpus, two
streams, plain device buffers standing in for the NVLink peer mappings -- ``PeerBuffers.local_group``), driven through
the C ABI the multi-GPU path uses -- ``tt_rescore_topk_fused`` (push) -> ``tt_merge_topk_fused`` (flag wait, merge,
margins, auto-merge) -> ``tt_exchange_push`` (second round) -> ``tt_peer_barrier`` -- and compared with the CPU oracle
on the whole corpus; then ``ShardedIndex`` itself (``search``, ``retrieve_host``, the second round after a repair) through
``LocalShardGroup``.  What a 1-GPU box cannot show is only the NVLink transport itself (tests/test_gpu_sharded.py)."""

import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def two_shards():
    from tensor_truth_b200 import _lib
    from tensor_truth_b200.index import DeviceIndex
    from tensor_truth_b200.sharded import shard_bounds
    from tensor_truth_b200.synth import make_small

    dev = torch.device("cuda:0")
    _lib.set_wait_timeout_ms(0, 1500)  # a protocol bug must fail the test (TTError), not hang the box
    tree, bits, inv, q = make_small(40_000, 24, dim=1024, levels=3, seed=77)
    shards = []
    for r in range(2):
        lo, hi = shard_bounds(bits.shape[0], 2, r)
        shards.append(DeviceIndex(bits[lo:hi], tree, id_base=lo, device=dev))
    yield tree, bits, q, shards
    _lib.set_wait_timeout_ms(0, 0)


def _merge(L, pb, lane, idx, b, k, out_scores, out_ids, all_margins=None, am=None, stream=None):
    from tensor_truth_b200 import _lib

    region = pb.region_ptr(lane)
    _lib.check(L.tt_merge_topk_fused(region, region + pb.ids_off, pb.world, pb.rec_stride // 4, pb.rec_stride // 8, b, k, k,
                                     idx.score_mode, out_scores.data_ptr(), out_ids.data_ptr(), C.byref(pb.desc(lane)),
                                     all_margins.data_ptr() if all_margins is not None else None,
                                     C.byref(am) if am is not None else None, stream))


@pytest.mark.parametrize("b,k", [(1, 10), (8, 10), (3, 100)])
def test_push_merge_automerge_matches_oracle(two_shards, b, k):
    import oracle
    from oracle import cport
    from tensor_truth_b200 import _lib
    from tensor_truth_b200.index import MergeResult
    from tensor_truth_b200.sharded import PeerBuffers

    tree, bits, q, shards = two_shards
    L = _lib.lib()
    dev = shards[0].device
    pbs = PeerBuffers.local_group(2, b, k, dev, lanes=2)
    streams = [torch.cuda.Stream(dev) for _ in range(2)]
    ids_o, sc_o, _ = cport.scan_topk(bits, q, k)
    outs = [(torch.empty((b, k), dtype=torch.float32, device=dev), torch.empty((b, k), dtype=torch.int64, device=dev))
            for _ in range(2)]
    mos = [MergeResult(torch.empty((b, 2 * k), dtype=torch.int64, device=dev), torch.empty((b, 2 * k), dtype=torch.float64, device=dev),
                       torch.empty((b,), dtype=torch.int32, device=dev)) for _ in range(2)]
    allm = [torch.zeros((2, b), dtype=torch.float32, device=dev) for _ in range(2)]
    margins = [torch.zeros((b,), dtype=torch.float32, device=dev) for _ in range(2)]
    torch.cuda.synchronize()
    # six epochs: both slots of a lane's ring several times over, alternating lanes like the two-stream pipeline does
    for step in range(6):
        lane = step % 2
        q0 = (step * b) % (q.shape[0] - b + 1)
        qd = torch.from_numpy(q[q0:q0 + b]).to(dev)
        torch.cuda.synchronize()
        for r in (1, 0):  # rank 1 first: rank 0's merge must really wait for a flag raised by another stream
            with torch.cuda.stream(streams[r]):
                # top-100 needs deeper per-CTA lists than the index default of 32 to be certifiable (the bench does the same)
                w = dict(shards[r]._buffers(b, k, slot=("t", lane), hi_only=False, kprime=128 if k > 32 else None))
                w["margin"] = margins[r]
                shards[r].search(qd, k, out=w, xchg=pbs[r].desc(lane))
        for r in (0, 1):
            with torch.cuda.stream(streams[r]):
                am = shards[r]._am_args(0.5, mos[r])
                _merge(L, pbs[r], lane, shards[r], b, k, outs[r][0], outs[r][1], allm[r], am, shards[r]._stream())
        torch.cuda.synchronize()
        _lib.check_status(0)
        for r in (0, 1):
            assert (outs[r][1].cpu().numpy() == ids_o[q0:q0 + b]).all(), (step, r)
            assert (outs[r][0].cpu().numpy() == sc_o[q0:q0 + b]).all(), (step, r)
            # every rank sees every rank's certificate margins
            assert torch.equal(allm[r][0], margins[0]) and torch.equal(allm[r][1], margins[1])
            assert bool((allm[r] > shards[r].eps).all())
            for i in range(b):
                exp = oracle.retrieve(bits, q[q0 + i], k, tree)
                n = int(mos[r].lens[i])
                got = [(int(o), float(s)) for o, s in zip(mos[r].ids[i, :n].tolist(), mos[r].scores[i, :n].tolist())]
                assert got == exp, (step, r, i)
    assert [int(e) for e in pbs[0].epochs.tolist()[:2]] == [3, 3]


def test_second_round_push_and_barrier(two_shards):
    """``tt_exchange_push`` of a finished record (what follows a host-side repair) + ``tt_peer_barrier``.

    Two simulated ranks share ONE GPU here, and two streams of one process may share a hardware queue: a kernel must
    never wait for one that was enqueued AFTER it (on real multi-GPU runs every rank has its own device, so the mutual
    wait of a barrier is fine there -- bench.py's timed loops start behind one).  So the barrier is exercised one
    rank at a time, the other rank's arrival being its real kernel run earlier (or, for the very first one, the flag
    store that kernel would have made)."""
    from oracle import cport
    from tensor_truth_b200 import _lib
    from tensor_truth_b200.sharded import PeerBuffers

    tree, bits, q, shards = two_shards
    L = _lib.lib()
    dev = shards[0].device
    b, k = 4, 10
    pbs = PeerBuffers.local_group(2, b, k, dev, lanes=1)
    streams = [torch.cuda.Stream(dev) for _ in range(2)]
    ids_o, sc_o, _ = cport.scan_topk(bits, q[:b], k)
    qd = torch.from_numpy(q[:b]).to(dev)
    outs = [(torch.empty((b, k), dtype=torch.float32, device=dev), torch.empty((b, k), dtype=torch.int64, device=dev))
            for _ in range(2)]

    def barrier_flags(r):  # int32 view of rank r's barrier flag array
        pb = pbs[r]
        off = pb.flags_off + pb.lanes * pb.SLOTS * pb.flag_stride * 4
        return pb.buf[off: off + 4 * pb.world].view(torch.int32)

    for rnd in range(3):
        for r in (0, 1):
            ex = shards[r].search_exact(qd, k)  # the "repaired" local record
            send, s_keys, s_ids, s_margins = pbs[r].send_record()
            s_keys.copy_(ex.keys)
            s_ids.copy_(ex.ids)
            s_margins.fill_(float("inf"))
        # barrier number rnd + 1: rank 1's arrival is simulated by the store its kernel makes into rank 0's flags ...
        barrier_flags(0)[1] = rnd + 1
        torch.cuda.synchronize()
        _lib.check(L.tt_peer_barrier(C.byref(pbs[0].desc(pbs[0].lanes)), shards[0]._stream()))
        torch.cuda.synchronize()
        _lib.check_status(0)
        assert barrier_flags(0).tolist() == [rnd + 1, rnd + 1] and int(barrier_flags(1)[0]) == rnd + 1  # rank 0 told everybody
        # ... and rank 1's real barrier kernel now finds rank 0's flag (and its own) in place
        _lib.check(L.tt_peer_barrier(C.byref(pbs[1].desc(pbs[1].lanes)), shards[1]._stream()))
        torch.cuda.synchronize()
        _lib.check_status(0)
        assert barrier_flags(1).tolist() == [rnd + 1, rnd + 1]
        for r in (1, 0):
            with torch.cuda.stream(streams[r]):
                _lib.check(L.tt_exchange_push(pbs[r].send_record()[0].data_ptr(), pbs[r].rec_bytes // 4 * 4,
                                              C.byref(pbs[r].desc(0)), shards[r]._stream()))
        for r in (0, 1):
            with torch.cuda.stream(streams[r]):
                _merge(L, pbs[r], 0, shards[r], b, k, outs[r][0], outs[r][1], None, None, shards[r]._stream())
        torch.cuda.synchronize()
        _lib.check_status(0)
        for r in (0, 1):
            assert (outs[r][1].cpu().numpy() == ids_o).all() and (outs[r][0].cpu().numpy() == sc_o).all(), (rnd, r)
    assert int(pbs[0].epochs[1]) == 3 and int(pbs[1].epochs[1]) == 3  # three barriers passed on both ranks
    assert int(pbs[0].epochs[0]) == 3 and int(pbs[1].epochs[0]) == 3  # three pushes on lane 0


def test_missing_peer_is_an_error_not_a_dead_context(two_shards):
    """A rank that never shows up: the waiting kernel gives up after the configured bound, the host raises TTError
    (TT_ERR_TIMEOUT) naming the rank -- and the CUDA context is still alive: the next query is answered."""
    from oracle import cport
    from tensor_truth_b200 import _lib
    from tensor_truth_b200.sharded import PeerBuffers

    tree, bits, q, shards = two_shards
    L = _lib.lib()
    dev = shards[0].device
    b, k = 1, 10
    _lib.set_wait_timeout_ms(0, 50)
    try:
        pbs = PeerBuffers.local_group(2, b, k, dev, lanes=1)
        qd = torch.from_numpy(q[:b]).to(dev)
        o = (torch.empty((b, k), dtype=torch.float32, device=dev), torch.empty((b, k), dtype=torch.int64, device=dev))
        shards[0].search(qd, k, xchg=pbs[0].desc(0))   # rank 0 pushes ...
        _merge(L, pbs[0], 0, shards[0], b, k, o[0], o[1], None, None, shards[0]._stream())  # ... and waits for a rank 1 that never comes
        torch.cuda.synchronize()
        with pytest.raises(_lib.TTError) as ei:
            _lib.check_status(0)
        assert ei.value.code == _lib.ERR_TIMEOUT and "rank 1" in str(ei.value)
        _lib.check_status(0)  # cleared
    finally:
        _lib.set_wait_timeout_ms(0, 1500)
    r = shards[0].search_certified(qd, k)
    torch.cuda.synchronize()
    lo = shards[0].id_base
    ids_o, sc_o, _ = cport.scan_topk(bits[lo:lo + shards[0].n_rows], q[:b], k)
    assert (r.ids.cpu().numpy() - lo == ids_o).all() and (r.scores.cpu().numpy() == sc_o).all()


def test_fused_tail_equals_the_two_launch_tail(two_shards):
    """``tt_rescore_topk_fused`` (one launch: re-score, last block selects + auto-merges) against ``tt_rescore_topk`` +
    ``tt_automerge`` (three launches) on the same shortlists: keys, scores, ids, margins and merged lists bit-equal."""
    from tensor_truth_b200 import _lib
    from tensor_truth_b200.index import MergeResult

    tree, bits, q, shards = two_shards
    idx = shards[0]
    L = _lib.lib()
    dev = idx.device
    for b, k in [(1, 10), (16, 10), (5, 37), (2, 200)]:
        qd = torch.from_numpy(q[:b]).to(dev)
        kp = 32 if k <= 16 else 128
        w = dict(idx._buffers(b, k, slot=("f", k), hi_only=False, kprime=kp))
        mo = MergeResult(torch.empty((b, 2 * k), dtype=torch.int64, device=dev), torch.empty((b, 2 * k), dtype=torch.float64, device=dev),
                         torch.empty((b,), dtype=torch.int32, device=dev))
        r = idx.search(qd, k, out=w, hi_only=False, am=idx._am_args(0.5, mo))
        torch.cuda.synchronize()
        n_cand = idx.n_lists * kp
        keys2 = torch.empty_like(r.keys)
        sc2 = torch.empty_like(r.scores)
        ids2 = torch.empty_like(r.ids)
        mg2 = torch.empty_like(r.margin)
        ws = torch.empty(int(L.tt_rescore_workspace_bytes(b, n_cand)), dtype=torch.uint8, device=dev)
        src = idx.corpus
        _lib.check(L.tt_rescore_topk(src.data_ptr(), _lib.DTYPE_BF16, idx.n_rows, idx.dim, idx.dim, idx.id_base, qd.data_ptr(), b,
                                     w["cand_ids"].data_ptr(), n_cand, w["cand_thresh"].data_ptr(), idx.n_lists, k, idx.score_mode,
                                     keys2.data_ptr(), sc2.data_ptr(), ids2.data_ptr(), mg2.data_ptr(), ws.data_ptr(), ws.numel(),
                                     idx._stream()))
        m2 = idx.automerge(ids2, sc2)
        torch.cuda.synchronize()
        assert torch.equal(keys2, r.keys) and torch.equal(sc2, r.scores) and torch.equal(ids2, r.ids) and torch.equal(mg2, r.margin)
        assert torch.equal(m2.lens, mo.lens) and torch.equal(m2.ids, mo.ids) and torch.equal(m2.scores, mo.scores)


def test_step_graph_replays_match_eager(two_shards):
    """``DeviceIndex.step_graph``: the 3-kernel pipeline replayed as a CUDA graph gives what the eager calls give."""
    import oracle

    tree, bits, q, shards = two_shards
    idx = shards[1]
    lo = idx.id_base
    g = idx.step_graph(1, 10)
    sub = bits[lo:lo + idx.n_rows]
    for i in range(5):
        g.q.copy_(torch.from_numpy(q[i:i + 1]))
        g.replay()
        torch.cuda.synchronize()
        assert float(g.result.margin[0]) > g.eps
        n = int(g.merged.lens[0])
        got = [(int(o), float(s)) for o, s in zip(g.merged.ids[0, :n].tolist(), g.merged.scores[0, :n].tolist())]
        leaf = oracle.exact_topk(sub, q[i:i + 1], 10, 0, id_base=lo)
        exp = oracle.auto_merge([(int(o), float(s)) for o, s in zip(leaf[0][0], leaf[1][0]) if o >= 0], tree.parent_of,
                                tree.child_count, tree.prev_id, tree.next_id)
        assert got == exp, i


# ---- ShardedIndex itself (not just the ABI calls under it) on one GPU: LocalShardGroup steps two simulated ranks in lockstep


def test_sharded_index_search_on_one_gpu(two_shards):
    import oracle
    from oracle import cport
    from tensor_truth_b200 import _lib
    from tensor_truth_b200.index import MergeResult
    from tensor_truth_b200.sharded import LocalShardGroup

    tree, bits, q, shards = two_shards
    dev = shards[0].device
    grp = LocalShardGroup(shards)
    assert all(rk.transport == "peer" and rk.peers(1, 10) is not None for rk in grp.ranks)  # no NCCL fallback underneath
    ids_o, sc_o, _ = cport.scan_topk(bits, q, 10)
    qd = torch.from_numpy(q).to(dev)
    for rep in range(5):  # both slots of the lane's ring several times over, two batch shapes (two PeerBuffers groups)
        for b in (1, 12):
            mos = [MergeResult(torch.empty((b, 20), dtype=torch.int64, device=dev), torch.empty((b, 20), dtype=torch.float64, device=dev),
                               torch.empty((b,), dtype=torch.int32, device=dev)) for _ in range(2)]
            res = grp.search(qd[rep:rep + b], 10, merged_out=mos)
            torch.cuda.synchronize()
            _lib.check_status(0)
            for r, (scores, ids) in enumerate(res):
                assert (ids.cpu().numpy() == ids_o[rep:rep + b]).all() and (scores.cpu().numpy() == sc_o[rep:rep + b]).all(), (rep, b, r)
                for i in range(b):
                    n = int(mos[r].lens[i])
                    got = [(int(o), float(s)) for o, s in zip(mos[r].ids[i, :n].tolist(), mos[r].scores[i, :n].tolist())]
                    assert got == oracle.retrieve(bits, q[rep + i], 10, tree), (rep, b, r, i)


def test_sharded_index_retrieve_host_on_one_gpu(two_shards):
    import oracle
    from tensor_truth_b200.sharded import LocalShardGroup

    tree, bits, q, shards = two_shards
    grp = LocalShardGroup(shards)
    for rep in range(4):
        for b in (1, 3):
            res = grp.retrieve_host(torch.from_numpy(q[rep:rep + b]), 10)
            for r, (ids_h, sc_h, lens) in enumerate(res):
                for i in range(b):
                    got = [(int(o), float(s)) for o, s in zip(ids_h[i, :lens[i]], sc_h[i, :lens[i]])]
                    assert got == oracle.retrieve(bits, q[rep + i], 10, tree), (rep, b, r, i)
    assert all(rk.second_rounds == 0 for rk in grp.ranks)
    with pytest.raises(RuntimeError):
        grp.ranks[0].step_graph(1, 10)  # graphs of the sharded step need one process per rank


def test_sharded_index_second_round_on_one_gpu():
    """One shard whose top-k cannot be proven (near-duplicate rows), one that can: EVERY rank must see it (the margins
    travel with the records), the rank concerned repairs, and all ranks exchange a second time."""
    import oracle
    from oracle import cport
    from tensor_truth_b200 import _lib
    from tensor_truth_b200.index import DeviceIndex
    from tensor_truth_b200.sharded import LocalShardGroup, shard_bounds
    from tensor_truth_b200.synth import make_small

    dev = torch.device("cuda:0")
    _lib.set_wait_timeout_ms(0, 1500)
    try:
        tree, bits, inv, q = make_small(20_000, 4, dim=1024, levels=3, seed=78)
        rng = np.random.default_rng(3)
        base = rng.standard_normal(1024).astype(np.float32)
        dup = oracle.f32_to_bf16_bits(base[None, :] * (1.0 + 2e-3 * rng.standard_normal((20_000, 1024)).astype(np.float32)))
        mixed = np.concatenate([bits[:20_000], dup])
        shards = []
        for r in range(2):
            lo, hi = shard_bounds(mixed.shape[0], 2, r)
            shards.append(DeviceIndex(mixed[lo:hi], None, id_base=lo, device=dev))
        grp = LocalShardGroup(shards)
        q2 = np.stack([q[0], (base * (1.0 + 1e-3 * rng.standard_normal(1024))).astype(np.float32), q[1]])
        ids_m, sc_m, _ = cport.scan_topk(mixed, q2, 10)
        for rep in range(3):
            res = grp.retrieve_host(torch.from_numpy(q2), 10, merge=False)
            for r, (ids_h, sc_h, lens) in enumerate(res):
                assert (ids_h == ids_m).all() and (sc_h == sc_m.astype(np.float64)).all(), (rep, r)
        assert [rk.second_rounds for rk in grp.ranks] == [3, 3]
        assert [rk.local.fallbacks > 0 for rk in grp.ranks] == [False, True]
    finally:
        _lib.set_wait_timeout_ms(0, 0)
