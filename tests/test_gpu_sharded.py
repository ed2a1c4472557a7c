"""Row-sharded path, one process per rank: ShardedIndex with the fused peer exchange and with the all-gather transport
must both reproduce the whole-corpus oracle answer bit-exactly.

With two GPUs (``gpurun --gpus 2``) the ranks sit on their own devices: NVLink peer mappings / NCCL.  On a ONE-GPU box
the all-gather transport still runs, as two processes sharing GPU 0 over the ``gloo`` backend (which moves CUDA tensors);
the peer transport's symmetric-memory rendezvous wants one device per rank, so that case is skipped here and covered by
tests/test_gpu_exchange_one_device.py (the ABI protocol and ``ShardedIndex`` itself through ``LocalShardGroup``)."""

import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, transport, out_dir, one_gpu=False):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import oracle
    from oracle import cport
    from tensor_truth_b200.index import DeviceIndex
    from tensor_truth_b200.sharded import ShardedIndex, shard_bounds
    from tensor_truth_b200.synth import make_small

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), TT_EXCHANGE=transport)
    if one_gpu:  # both ranks on GPU 0, collectives through gloo (the all-gather transport only)
        assert transport == "nccl"
        torch.cuda.set_device(0)
        dev = torch.device("cuda:0")
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        torch.cuda.set_device(rank)
        dev = torch.device(f"cuda:{rank}")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    tree, bits, inv, q = make_small(50_000, 300, dim=1024, levels=3, seed=31)
    lo, hi = shard_bounds(bits.shape[0], world, rank)
    idx = DeviceIndex(bits[lo:hi], tree, id_base=lo, device=dev)
    sh = ShardedIndex(idx)
    assert sh.transport == ("peer" if transport == "peer" else "nccl")
    ids_o, sc_o, _ = cport.scan_topk(bits, q, 10)
    qd = torch.from_numpy(q).to(dev)
    for rep in range(6):  # several epochs through the slot ring, two batch shapes
        for b in (1, 12):
            scores, ids = sh.search(qd[:b], 10)
            torch.cuda.synchronize()
            assert (ids.cpu().numpy() == ids_o[:b]).all() and (scores.cpu().numpy() == sc_o[:b]).all(), (rep, b)
    # a wide batch: every rank's stage 1 is the GEMM-shaped scan (one shortlist per query), same exchange
    assert idx._use_gemm(300)
    for rep in range(2):
        scores, ids = sh.search(qd, 10)
        torch.cuda.synchronize()
        assert (ids.cpu().numpy() == ids_o).all() and (scores.cpu().numpy() == sc_o).all(), rep
    if transport == "peer":
        # the whole sharded step as ONE CUDA graph per lane (device-resident exchange epochs): two lanes on two streams,
        # replayed alternately like the bench's throughput loop, every replay checked against the oracle
        lanes = [sh.step_graph(1, 10, 0.5, lane) for lane in range(2)]
        streams = [torch.cuda.Stream(dev) for _ in range(2)]
        torch.cuda.synchronize()
        kept = []
        for step in range(10):
            lane = step % 2
            g = lanes[lane]
            with torch.cuda.stream(streams[lane]):
                g.q.copy_(qd[step:step + 1], non_blocking=True)
                g.replay()
                kept.append((step, g.result.ids.clone(), g.result.scores.clone(), g.merged.ids.clone(), g.merged.scores.clone(),
                             g.merged.lens.clone(), g.result.margin.clone()))
        torch.cuda.synchronize()
        from tensor_truth_b200 import _lib as _l

        _l.check_status(rank)
        for step, ids, scores, mids, msc, mlens, margin in kept:
            assert (ids.cpu().numpy() == ids_o[step:step + 1]).all() and (scores.cpu().numpy() == sc_o[step:step + 1]).all(), step
            assert float(margin[0]) > lanes[0].eps
            exp = oracle.retrieve(bits, q[step], 10, tree)
            n = int(mlens[0])
            assert [(int(o), float(s)) for o, s in zip(mids[0, :n].tolist(), msc[0, :n].tolist())] == exp, step
        # the host path switches to its graph after GRAPH_AFTER eager calls: results must not change across the switch
        for rep in range(6):
            ids_h, sc_h, lens = sh.retrieve_host(torch.from_numpy(q[rep:rep + 1]), 10)
            exp = oracle.retrieve(bits, q[rep], 10, tree)
            assert [(int(o), float(s)) for o, s in zip(ids_h[0, :lens[0]], sc_h[0, :lens[0]])] == exp, rep
    if transport == "peer":
        # two serving threads per rank, thread t on lane t of every rank (different query sequences per lane), past the
        # point where each lane captures its graph: the lanes' exchanges interleave freely, every answer is the oracle's
        import threading

        errs = []

        def serve(t):
            try:
                torch.cuda.set_device(rank)
                for i in range(10):
                    qi = (7 * t + 3 * i) % 64
                    ids_t, sc_t, lens_t = sh.retrieve_host(torch.from_numpy(q[qi:qi + 1]), 10, lane=t)
                    exp_t = oracle.retrieve(bits, q[qi], 10, tree)
                    assert [(int(o), float(s)) for o, s in zip(ids_t[0, :lens_t[0]], sc_t[0, :lens_t[0]])] == exp_t, (t, i)
            except BaseException as exc:  # noqa: BLE001
                errs.append(exc)

        threads = [threading.Thread(target=serve, args=(t,)) for t in range(2)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        if errs:
            raise errs[0]
    ids_h, sc_h, lens = sh.retrieve_host(torch.from_numpy(q[:3]), 10)
    for b in range(3):
        exp = oracle.retrieve(bits, q[b], 10, tree)
        got = [(int(o), float(s)) for o, s in zip(ids_h[b, :lens[b]], sc_h[b, :lens[b]])]
        assert got == exp
    assert sh.transport == ("peer" if transport == "peer" else "nccl")  # no silent fallback
    # one shard whose top-k cannot be proven (near-duplicate rows), one that can: the host path must notice on EVERY rank
    # (the margins travel with the records), repair on the rank concerned and exchange a second time
    rng = np.random.default_rng(3)
    base = rng.standard_normal(1024).astype(np.float32)
    dup = oracle.f32_to_bf16_bits(base[None, :] * (1.0 + 2e-3 * rng.standard_normal((25_000, 1024)).astype(np.float32)))
    mixed = np.concatenate([bits[:25_000], dup])
    lo2, hi2 = shard_bounds(mixed.shape[0], world, rank)
    sh2 = ShardedIndex(DeviceIndex(mixed[lo2:hi2], None, id_base=lo2, device=dev))
    q2 = np.stack([q[0], (base * (1.0 + 1e-3 * rng.standard_normal(1024))).astype(np.float32), q[1]])
    ids_m, sc_m, _ = cport.scan_topk(mixed, q2, 10)
    for rep in range(3):
        ids_h, sc_h, lens = sh2.retrieve_host(torch.from_numpy(q2), 10, merge=False)
        assert (ids_h == ids_m).all() and (sc_h == sc_m.astype(np.float64)).all(), rep
    if transport == "peer":
        assert sh2.second_rounds == 3      # every rank took the second round, every time
        # the same hard batch from two serving threads at once (lane t each): both lanes take second rounds concurrently --
        # per-lane send records and rings, the repair ladder's shared workspaces behind its lock -- and stay exact
        import threading

        errs2 = []

        def serve_hard(t):
            try:
                torch.cuda.set_device(rank)
                for rep in range(3):
                    ids_t, sc_t, _ = sh2.retrieve_host(torch.from_numpy(q2), 10, merge=False, lane=t)
                    assert (ids_t == ids_m).all() and (sc_t == sc_m.astype(np.float64)).all(), (t, rep)
            except BaseException as exc:  # noqa: BLE001
                errs2.append(exc)

        threads = [threading.Thread(target=serve_hard, args=(t,)) for t in range(2)]
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        if errs2:
            raise errs2[0]
        assert sh2.second_rounds == 9
    assert (sh2.local.fallbacks > 0) == (rank == 1)
    open(os.path.join(out_dir, f"ok-{transport}-{rank}"), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.parametrize("transport", ["peer", "nccl"])
def test_sharded_index_two_ranks(tmp_path, transport):
    import torch.multiprocessing as mp

    one_gpu = torch.cuda.device_count() < 2
    if one_gpu and transport == "peer":
        pytest.skip("the symmetric-memory rendezvous wants one GPU per rank; on one GPU the peer protocol and ShardedIndex "
                    "are covered by tests/test_gpu_exchange_one_device.py (LocalShardGroup)")
    mp.spawn(_worker, args=(2, _free_port(), transport, str(tmp_path), one_gpu), nprocs=2, join=True)
    assert all(os.path.exists(tmp_path / f"ok-{transport}-{r}") for r in range(2))
