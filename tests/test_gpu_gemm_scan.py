"""Parity of the GEMM-shaped stage 1 (scan_gemm.cu: wide batches, the tensor-bound regime of BASELINE configs[3])
against the CPU oracle, through the C ABI (tt_scan_gemm_topk_bf16) and through DeviceIndex.search.

ids / keys / scores of the final top-k are bit-exact; the stage-1 shortlist is checked against its contract
(every row it leaves out scores <= out_thresh, within the stated stage-1 error bound).

Every test here needs a GPU:  python -m pytest tests -m gpu
"""

import numpy as np
import pytest
import torch

import oracle
from oracle import cport
from tensor_truth_b200 import _lib
from tensor_truth_b200 import index as index_mod
from tensor_truth_b200.synth import SynthCorpus, make_small

pytestmark = pytest.mark.gpu

EPS_HI = index_mod.EPS_BF16_CORPUS + index_mod.EPS_HI_ONLY


def _index(bits, tree=None, **kw):
    return index_mod.DeviceIndex(bits, tree, device=torch.device("cuda:0"), **kw)


def _np(t):
    return t.detach().cpu().numpy()


def _gemm_scan(idx, q, kp):
    """One direct call of the C-ABI entry point: (ids [B,kp], approx [B,kp], thresh [B])."""
    L, ptr = idx.lib, _lib.ptr
    b = int(q.shape[0])
    dev = idx.device
    q_hi = torch.empty((b, idx.dim), dtype=torch.bfloat16, device=dev)
    ids = torch.empty((b, kp), dtype=torch.int64, device=dev)
    approx = torch.empty((b, kp), dtype=torch.float32, device=dev)
    thresh = torch.empty((b,), dtype=torch.float32, device=dev)
    ws = torch.empty(int(L.tt_scan_gemm_workspace_bytes(b, kp)), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.tt_prepare_queries(ptr(q), b, idx.dim, ptr(q_hi), None, st))
    _lib.check(L.tt_scan_gemm_topk_bf16(ptr(idx.corpus), idx.n_rows, idx.dim, idx.dim, ptr(idx.inv_norm), ptr(q_hi), None, b, kp,
                                        idx.id_base, ptr(ids), ptr(approx), ptr(thresh), ptr(ws), ws.numel(), st))
    torch.cuda.synchronize()
    return _np(ids), _np(approx), _np(thresh)


def _exact_cosines(bits, q):
    c = oracle.bf16_bits_to_f32(bits).astype(np.float64)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    qq = q.astype(np.float64)
    qq /= np.linalg.norm(qq, axis=1, keepdims=True)
    return qq @ c.T  # [B, n_rows]


@pytest.fixture(scope="module")
def mid():
    tree, bits, inv, q = make_small(30_000, 300, dim=1024, levels=3, seed=31)
    return tree, bits, inv, q


@pytest.mark.parametrize("kp", [128, 512])
def test_shortlist_contract(mid, kp):
    """out_ids are the K' best rows by approximate score, sorted; out_thresh bounds every row left out."""
    tree, bits, inv, q = mid
    idx = _index(bits, tree)
    ids, approx, thresh = _gemm_scan(idx, torch.from_numpy(q).cuda(), kp)
    exact = _exact_cosines(bits, q)
    n = bits.shape[0]
    assert (ids >= 0).all() and (ids < n).all()
    for b in range(q.shape[0]):
        assert len(set(ids[b].tolist())) == kp                      # no row twice
        assert (np.diff(approx[b]) <= 0).all()                      # best first
        assert thresh[b] == approx[b, kp - 1]                       # the K'-th score is the bound on what was dropped
        err = np.abs(exact[b, ids[b]] - approx[b]).max()
        assert err < EPS_HI / 4, err                                # stage-1 error well inside the certificate bound
        out = np.ones(n, bool)
        out[ids[b]] = False
        assert exact[b, out].max() <= thresh[b] + EPS_HI / 4        # nothing better was left out


@pytest.mark.parametrize("k", [10, 100])
def test_wide_batch_topk_bit_exact(mid, k):
    tree, bits, inv, q = mid
    ids_o, sc_o, keys_o = cport.scan_topk(bits, q, k)
    idx = _index(bits, tree)
    qd = torch.from_numpy(q).cuda()
    assert idx._use_gemm(int(qd.shape[0]))
    r = idx.search(qd, k)
    torch.cuda.synchronize()
    proven = _np(r.margin) > r.eps
    assert proven.mean() > 0.9, _np(r.margin)
    assert (_np(r.ids)[proven] == ids_o[proven]).all() and (_np(r.scores)[proven] == sc_o[proven]).all()
    r = idx.search_certified(qd, k)
    torch.cuda.synchronize()
    assert (_np(r.ids) == ids_o).all() and (_np(r.scores) == sc_o).all() and (_np(r.keys) == keys_o).all()
    assert idx.fallbacks == 0
    # the whole path: auto-merge on the GEMM path's top-k == the oracle's retrieve()
    m = idx.automerge(r.ids, r.scores)
    torch.cuda.synchronize()
    mi, ms, ml = _np(m.ids), _np(m.scores), _np(m.lens)
    for b in (0, 131, 299):
        exp = oracle.retrieve(bits, q[b], k, tree)
        assert [(int(o), float(s)) for o, s in zip(mi[b, :ml[b]], ms[b, :ml[b]])] == exp


@pytest.mark.parametrize("n_rows", [1, 100, 255, 256, 257, 20_001])
def test_ragged_rows_and_query_blocks(n_rows):
    """Row counts around the 256-row super-tile, fewer rows than K', and a query count that leaves a 1-query block."""
    rng = np.random.default_rng(n_rows)
    c = rng.standard_normal((n_rows, 128)).astype(np.float32)
    if n_rows > 4:
        c[3] = c[1]  # exact duplicates: ties broken by the smaller id
        c[n_rows - 1] = c[1]
    bits = oracle.f32_to_bf16_bits(c)
    q = rng.standard_normal((257, 128)).astype(np.float32)
    q[256] = oracle.bf16_bits_to_f32(bits[min(1, n_rows - 1)])  # the 1-query block asks for a stored (duplicated) row
    idx = _index(bits, None)
    ids, approx, thresh = _gemm_scan(idx, torch.from_numpy(q).cuda(), 128)
    if n_rows <= 128:
        assert np.isneginf(thresh).all()                            # nothing was left out
        assert ((ids >= 0).sum(axis=1) == n_rows).all()
        assert all(sorted(row[:n_rows].tolist()) == list(range(n_rows)) for row in ids)
        assert (ids[:, n_rows:] == -1).all() and np.isneginf(approx[:, n_rows:]).all()
    ids_o, sc_o, _ = cport.scan_topk(bits, q, 10)
    r = idx.search_certified(torch.from_numpy(q).cuda(), 10)
    torch.cuda.synchronize()
    assert (_np(r.ids) == ids_o).all() and (_np(r.scores) == sc_o).all()
    if n_rows > 4:
        assert _np(r.ids)[256, :3].tolist() == [1, 3, n_rows - 1]


def test_id_base_and_slices(mid, monkeypatch):
    """A shard (id_base > 0) searched in GEMM slices smaller than the batch gives the same answer as one slice."""
    tree, bits, inv, q = mid
    ids_o, sc_o, _ = cport.scan_topk(bits[10_000:], q, 10)
    qd = torch.from_numpy(q).cuda()
    for sl in ("4096", "256"):
        monkeypatch.setenv("TT_GEMM_SLICE", sl)
        idx = _index(bits[10_000:], None, id_base=10_000)
        r = idx.search_certified(qd, 10)
        torch.cuda.synchronize()
        assert (_np(r.ids) == ids_o + 10_000).all() and (_np(r.scores) == sc_o).all()


def test_buffer_overflow_is_reported_and_repaired(mid, monkeypatch):
    """A phase that appends more than the buffer holds must say so (thresh = +inf): the certificate refuses and the
    repair ladder (hi+lo re-scan) answers -- never a silently truncated shortlist."""
    tree, bits, inv, q = mid
    ids_o, sc_o, _ = cport.scan_topk(bits, q, 10)
    monkeypatch.setenv("TT_GEMM_GROWTH", "60")  # second phase visits 60x the sample: ~60 K' appends > 16 K' capacity
    idx = _index(bits, None)
    _, _, thresh = _gemm_scan(idx, torch.from_numpy(q).cuda(), 128)
    assert np.isposinf(thresh).mean() > 0.5
    r = idx.search_certified(torch.from_numpy(q).cuda(), 10)
    torch.cuda.synchronize()
    assert idx.retries > 0
    assert (_np(r.ids) == ids_o).all() and (_np(r.scores) == sc_o).all()


def test_near_duplicates_in_a_wide_batch():
    """Rows the approximate scores cannot separate: every certificate refuses, the ladder ends in the exact scan."""
    rng = np.random.default_rng(23)
    base = rng.standard_normal(256).astype(np.float32)
    c = base[None, :] * (1.0 + 2e-3 * rng.standard_normal((20_000, 256)).astype(np.float32))
    c[777] = c[12]
    bits = oracle.f32_to_bf16_bits(c)
    q = (base[None, :] * (1.0 + 1e-3 * rng.standard_normal((260, 256)))).astype(np.float32)
    ids_o, sc_o, _ = cport.scan_topk(bits, q, 10)
    idx = _index(bits, None)
    r = idx.search_certified(torch.from_numpy(q).cuda(), 10)
    torch.cuda.synchronize()
    assert idx.fallbacks > 0
    assert (_np(r.ids) == ids_o).all() and (_np(r.scores) == sc_o).all()


def test_big_gemm_path_equals_pair_kernel_path_and_exact_scan(monkeypatch):
    """2M x 1024, 1024 queries, top-100: the GEMM-shaped scan, the 64-query CTA-pair scan and (on a few queries) the
    exact fp64 scan agree bit for bit -- the size-independent check for configs[3]."""
    n = 2_000_000
    sc = SynthCorpus(n, 1024, levels=3, seed=77, device="cuda")
    corpus, inv = sc.rows(0, n)
    q = sc.finish_queries(sc.queries(1024, lookup=lambda t: corpus[t])).cuda()
    idx = _index(corpus, sc.tree, inv_norm=inv)
    r = idx.search_certified(q, 100)
    torch.cuda.synchronize()
    g_ids, g_sc = r.ids.clone(), r.scores.clone()
    assert idx.fallbacks == 0
    monkeypatch.setenv("TT_NO_GEMM", "1")
    idx2 = _index(corpus, sc.tree, inv_norm=inv)
    assert not idx2._use_gemm(1024)
    r2 = idx2.search_certified(q, 100)
    torch.cuda.synchronize()
    assert torch.equal(g_ids, r2.ids) and torch.equal(g_sc, r2.scores)
    ex = idx.search_exact(q[:4], 100)
    torch.cuda.synchronize()
    assert torch.equal(g_ids[:4], ex.ids) and torch.equal(g_sc[:4], ex.scores)


@pytest.mark.parametrize("b", [1, 40, 300])
def test_massive_ties_resolve_by_smaller_id(b):
    """A corpus made of verbatim copies of 40 distinct vectors: every score is shared by ~750 rows, so the top-k is
    decided by the tie rule alone (smaller id first).  No approximate shortlist can prove that; whatever the ladder
    does (batch-1 kernel, GEMM-shaped scan, exact scan), the answer must be the oracle's."""
    rng = np.random.default_rng(41)
    base = rng.standard_normal((40, 256)).astype(np.float32)
    assign = rng.integers(0, 40, size=30_000)
    bits = oracle.f32_to_bf16_bits(base[assign])
    q = (base[rng.integers(0, 40, size=b)] + 0.05 * rng.standard_normal((b, 256))).astype(np.float32)
    ids_o, sc_o, _ = cport.scan_topk(bits, q, 10)
    idx = _index(bits, None)
    r = idx.search_certified(torch.from_numpy(q).cuda(), 10)
    torch.cuda.synchronize()
    assert (_np(r.ids) == ids_o).all() and (_np(r.scores) == sc_o).all()
    assert (np.diff(_np(r.ids), axis=1) > 0).all()  # ten copies of the best vector, ids ascending
    if b > 32:  # one shortlist of 128 per query cannot hold ~750 tied rows: the certificate must have refused
        assert idx.retries + idx.fallbacks > 0


@pytest.mark.parametrize("b", [17, 24, 32])
def test_hilo_pass_17_to_32_queries(mid, b):
    """Batches of 17-32 hi+lo queries take the 64-column GEMM-shaped pass with (hi, lo) column pairs: the approximate
    scores carry 16 mantissa bits of the query (error far below the hi+lo certificate bound), the shortlist contract
    holds, and the certified result equals the oracle's bit for bit with no repair."""
    tree, bits, inv, q = mid
    idx = _index(bits, tree)
    assert idx._use_gemm_hilo(b)
    qd = torch.from_numpy(q[:b]).cuda()
    # the C-ABI entry point directly
    L, ptr = idx.lib, _lib.ptr
    kp = 128
    q_hi = torch.empty((b, idx.dim), dtype=torch.bfloat16, device=idx.device)
    q_lo = torch.empty_like(q_hi)
    ids = torch.empty((b, kp), dtype=torch.int64, device=idx.device)
    approx = torch.empty((b, kp), dtype=torch.float32, device=idx.device)
    thresh = torch.empty((b,), dtype=torch.float32, device=idx.device)
    ws = torch.empty(int(L.tt_scan_gemm_workspace_bytes(b, kp)), dtype=torch.uint8, device=idx.device)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.tt_prepare_queries(ptr(qd), b, idx.dim, ptr(q_hi), ptr(q_lo), st))
    _lib.check(L.tt_scan_gemm_topk_bf16(ptr(idx.corpus), idx.n_rows, idx.dim, idx.dim, ptr(idx.inv_norm), ptr(q_hi), ptr(q_lo), b, kp,
                                        idx.id_base, ptr(ids), ptr(approx), ptr(thresh), ptr(ws), ws.numel(), st))
    torch.cuda.synchronize()
    cos = _exact_cosines(bits, q[:b])
    ids_h, approx_h, thresh_h = _np(ids), _np(approx), _np(thresh)
    for i in range(b):
        assert (ids_h[i] >= 0).all() and len(set(ids_h[i].tolist())) == kp
        err = np.abs(approx_h[i] - cos[i, ids_h[i]]).max()
        assert err < 2.5e-4 / 4, err                         # hi+lo precision, not hi-only (which is ~1e-3)
        left_out = np.ones(bits.shape[0], bool)
        left_out[ids_h[i]] = False
        assert cos[i, left_out].max() <= thresh_h[i] + 2.5e-4  # the out_thresh contract
    # through the index: certified, no retries
    idx.retries = idx.deep_rescans = idx.fallbacks = 0
    r = idx.search_certified(qd, 10)
    torch.cuda.synchronize()
    ids_o, sc_o, _ = cport.scan_topk(bits, q[:b], 10)
    assert (_np(r.ids) == ids_o).all() and (_np(r.scores) == sc_o).all()
    assert r.eps < 1e-3 and not r.hi_only and idx.retries == idx.deep_rescans == idx.fallbacks == 0


def test_hi_only_certificate_credit(mid, monkeypatch):
    """A hi-only scan's error bound budgets 2^-8 for the part of the query the tensor cores never see; the actual
    rho = |q/|q| - bf16(q/|q|)| is measured per query (tt_prepare_queries_rho) and the difference is handed to the
    certificate by lowering the query's thresholds (tt_certificate_credit).  Checked: rho against numpy (an upper bound,
    and a tight one), the margins grow by exactly eps_hi_only - rho, the true score error of every row stays below
    eps_bf16 + rho (what the credited certificate assumes), and the certified answers are the exact ones."""
    tree, bits, inv, q = mid
    idx = _index(bits, tree)
    L, ptr = idx.lib, _lib.ptr
    b = 96
    qd = torch.from_numpy(q[:b]).cuda()
    q_hi = torch.empty((b, idx.dim), dtype=torch.bfloat16, device="cuda")
    rho = torch.empty((b,), dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.tt_prepare_queries_rho(ptr(qd), b, idx.dim, ptr(q_hi), None, ptr(rho), st))
    torch.cuda.synchronize()
    qn = q[:b].astype(np.float64)
    qn /= np.linalg.norm(qn, axis=1, keepdims=True)
    hi = q_hi.float().cpu().numpy().astype(np.float64)
    true_rho = np.linalg.norm(qn - hi, axis=1)
    got = _np(rho).astype(np.float64)
    assert (got >= true_rho).all() and (got <= true_rho * 1.002 + 3e-6).all()
    assert got.max() < index_mod.EPS_HI_ONLY and got.mean() < 0.6 * index_mod.EPS_HI_ONLY  # the point of measuring it
    # the hi-only approximate score of EVERY row is within eps_bf16 + rho of the exact cosine
    cos = _exact_cosines(bits, q[:b])
    c = oracle.bf16_bits_to_f32(bits).astype(np.float64)
    approx_all = (hi @ c.T) / np.linalg.norm(c, axis=1)[None, :]
    assert (np.abs(approx_all - cos).max(axis=1) <= index_mod.EPS_BF16_CORPUS + got).all()
    # margins with and without the credit
    r1 = idx.search(qd, 10)
    m1 = _np(r1.margin).copy()
    ids1 = _np(r1.ids).copy()
    monkeypatch.setenv("TT_NO_CERT_CREDIT", "1")
    r0 = idx.search(qd, 10)
    m0 = _np(r0.margin).copy()
    monkeypatch.delenv("TT_NO_CERT_CREDIT")
    assert r1.eps == r0.eps == pytest.approx(EPS_HI)
    np.testing.assert_allclose(m1 - m0, index_mod.EPS_HI_ONLY - got, rtol=0, atol=2e-6)
    ids_o, sc_o, _ = cport.scan_topk(bits, q[:b], 10)
    proven = m1 > r1.eps
    assert proven.all() and (ids1 == ids_o).all()
    # the unit of the credit kernel itself: -inf / +inf thresholds keep their meaning, finite ones drop by the credit
    th = torch.tensor([[0.5, -float("inf")], [float("inf"), 0.25]], dtype=torch.float32, device="cuda")
    rr = torch.tensor([0.001, 0.5], dtype=torch.float32, device="cuda")  # second query: rho > eps -> no credit
    _lib.check(L.tt_certificate_credit(ptr(th), 2, 2, ptr(rr), 0.004, st))
    torch.cuda.synchronize()
    out = _np(th)
    assert out[0, 0] == pytest.approx(0.497, abs=1e-6) and np.isneginf(out[0, 1]) and np.isposinf(out[1, 0]) and out[1, 1] == 0.25
