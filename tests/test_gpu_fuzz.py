"""Randomised end-to-end parity: random shapes (rows, width, k, batch, store dtype, score mode, id base, clustered or iid
rows, injected duplicates) through DeviceIndex.search_certified -- whichever stage-1 kernel the shape selects and whatever
the repair ladder has to do -- against the strict C oracle.  ids, keys and scores bit-exact.

Every test here needs a GPU:  python -m pytest tests -m gpu
"""

import numpy as np
import pytest
import torch

import oracle
from oracle import cport
from tensor_truth_b200 import _lib
from tensor_truth_b200.index import DeviceIndex

pytestmark = pytest.mark.gpu


def _case(seed):
    rng = np.random.default_rng(1000 + seed)
    dim = int(rng.choice([128, 256, 384, 1024]))
    n = int(rng.choice([1, 9, 130, 1000, 5000, 20_000, 47_111]))
    b = int(rng.choice([1, 3, 17, 33, 40, 64, 70, 130, 300]))
    k = int(rng.choice([1, 5, 10, 32, 33, 64, 100]))
    fp32_store = bool(rng.integers(0, 2))
    mode = int(rng.integers(0, 2))
    id_base = int(rng.choice([0, 12_345, 3_000_000_000]))
    if rng.integers(0, 2):  # clustered rows: neighbours are near-duplicates of a few centres
        centres = rng.standard_normal((max(1, n // 50), dim)).astype(np.float32)
        c = centres[rng.integers(0, centres.shape[0], size=n)] + 0.3 * rng.standard_normal((n, dim)).astype(np.float32)
    else:
        c = rng.standard_normal((n, dim)).astype(np.float32)
    if mode == 1:  # chroma_l2_exp: half of the cases with unit rows (certified through the norm bounds), half without (exact scan)
        if rng.integers(0, 2):
            c /= np.linalg.norm(c, axis=1, keepdims=True)
    for _ in range(min(n // 3, 20)):  # verbatim duplicates: ties by id
        c[rng.integers(0, n)] = c[rng.integers(0, n)]
    stored = c if fp32_store else oracle.f32_to_bf16_bits(c)
    q = rng.standard_normal((b, dim)).astype(np.float32)
    take = rng.integers(0, n, size=b)
    near = rng.random(b) < 0.5
    base_rows = c if fp32_store else oracle.bf16_bits_to_f32(stored)
    q[near] = base_rows[take[near]] + 0.1 * rng.standard_normal((int(near.sum()), dim)).astype(np.float32)
    return dict(dim=dim, n=n, b=b, k=k, fp32=fp32_store, mode=mode, id_base=id_base), stored, q


@pytest.mark.parametrize("seed", range(36))
def test_random_shapes_against_the_strict_oracle(seed):
    cfg, stored, q = _case(seed)
    ids_o, sc_o, keys_o = cport.scan_topk(stored, q, cfg["k"], cfg["mode"], cfg["id_base"])
    idx = DeviceIndex(stored, None, device=torch.device("cuda:0"), score_mode=cfg["mode"], id_base=cfg["id_base"])
    r = idx.search_certified(torch.from_numpy(q).cuda(), cfg["k"])
    torch.cuda.synchronize()
    g_ids, g_sc, g_keys = (t.cpu().numpy() for t in (r.ids, r.scores, r.keys))
    assert (g_ids == ids_o).all(), cfg
    assert (g_keys == keys_o).all(), cfg
    assert (g_sc == sc_o).all(), cfg


@pytest.mark.parametrize("dim", [8, 72, 96, 200])
@pytest.mark.parametrize("b", [1, 40, 130])
def test_widths_the_tensor_core_kernels_do_not_take(dim, b):
    """Embedding widths that are not multiples of 64 / 128 go through the CUDA-core scan, whatever the batch size."""
    rng = np.random.default_rng(dim * 7 + b)
    c = rng.standard_normal((6000, dim)).astype(np.float32)
    c[17] = c[4000]
    bits = oracle.f32_to_bf16_bits(c)
    q = rng.standard_normal((b, dim)).astype(np.float32)
    q[0] = oracle.bf16_bits_to_f32(bits[17])
    ids_o, sc_o, keys_o = cport.scan_topk(bits, q, 7)
    idx = DeviceIndex(bits, None, device=torch.device("cuda:0"))
    assert not idx._use_gemm(b) or dim % 64 == 0
    r = idx.search_certified(torch.from_numpy(q).cuda(), 7)
    torch.cuda.synchronize()
    assert (r.ids.cpu().numpy() == ids_o).all() and (r.scores.cpu().numpy() == sc_o).all()
    assert r.ids[0, :2].tolist() == [17, 4000]
