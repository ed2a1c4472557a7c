"""Host-side logic of the row-sharded path on CPU: two processes over gloo, the record layout, the
all-gather and the merge contract.  The per-rank top-k and the merge come from the oracle here (the
CUDA kernels that normally sit in those two slots are covered by the -m gpu tests)."""

import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import oracle
    from tensor_truth_b200.sharded import ShardedSearch, record_layout, shard_bounds

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(os.path.join(ROOT, "tests", "golden", "mini_scan.npz"))
    bits, q = g["bits"], g["queries"]
    n = bits.shape[0]
    lo, hi = shard_bounds(n, world, rank, align=128)
    assert shard_bounds(n, world, 0)[0] == 0 and shard_bounds(n, world, world - 1)[1] == n

    def local_search(qt, k, keys_out, ids_out, slot=0):
        ids, _sc, keys = oracle.exact_topk(bits[lo:hi], qt.numpy(), k, 0, id_base=lo)
        keys_out.copy_(torch.from_numpy(keys))
        ids_out.copy_(torch.from_numpy(ids))

    def merge(recv, w, b, k, k_out, slot=0):
        rec, ids_off, _ = record_layout(b, k)
        assert recv.shape == (w, rec)
        keys = [recv[r, : b * k * 4].view(torch.float32).view(b, k).numpy() for r in range(w)]
        ids = [recv[r, ids_off: ids_off + b * k * 8].view(torch.int64).view(b, k).numpy() for r in range(w)]
        o_s = np.empty((b, k_out), np.float32)
        o_i = np.empty((b, k_out), np.int64)
        for i in range(b):
            kk, ii = oracle.merge_topk_lists([x[i] for x in keys], [x[i] for x in ids], k_out)
            o_s[i], o_i[i] = kk, ii
        return torch.from_numpy(o_s), torch.from_numpy(o_i)

    ss = ShardedSearch(local_search, merge, torch.device("cpu"))
    assert ss.world == world and ss.rank == rank
    for k in (10, 37):
        scores, ids = ss.search(torch.from_numpy(q), k)
        assert (ids.numpy() == g[f"cos_k{k}_ids"]).all()
        assert (scores.numpy() == g[f"cos_k{k}_scores"]).all()
    # odd batch: the keys section is padded to 8 bytes so the ids stay aligned
    rec, ids_off, _ = record_layout(3, 1)
    assert ids_off % 8 == 0 and rec == ids_off + 3 * 8
    scores, ids = ss.search(torch.from_numpy(q[:3]), 1)
    assert (ids.numpy()[:, 0] == g["cos_k10_ids"][:3, 0]).all()
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_sharded_search_two_ranks_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_shard_bounds_cover_and_align():
    from tensor_truth_b200.sharded import shard_bounds

    for n in (0, 1, 127, 128, 1000, 10_000_000, 100_000_000):
        for w in (1, 2, 4, 8):
            cuts = [shard_bounds(n, w, r) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            for (a, b), (c, d) in zip(cuts[:-1], cuts[1:]):
                assert b == c and a <= b
                assert b % 128 == 0 or b == n
