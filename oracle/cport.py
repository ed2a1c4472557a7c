"""ctypes binding of the C restatement (``oracle/liboracle.so``; build with ``make -C oracle``).

TEST/BENCH INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        L.oracle_scan_topk.restype = C.c_int
        L.oracle_scan_topk.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_int,
                                       C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_automerge.restype = C.c_int
        L.oracle_automerge.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_fast_scan_topk.restype = C.c_int
        L.oracle_fast_scan_topk.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int,
                                            C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
        L.oracle_fast_threads.restype = C.c_int
        L.oracle_fast_set_threads.restype = None
        L.oracle_fast_set_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def scan_topk(corpus: np.ndarray, queries: np.ndarray, k: int, score_mode: int = 0, id_base: int = 0):
    """Strict C oracle.  ``corpus`` uint16 (bf16 bits) or float32 ``[N, D]``."""
    corpus = np.ascontiguousarray(corpus)
    dtype = 0 if corpus.dtype == np.uint16 else 1
    if dtype == 1:
        corpus = corpus.astype(np.float32, copy=False)
    q = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
    n, d = corpus.shape
    b = q.shape[0]
    scores = np.empty((b, k), np.float32)
    keys = np.empty((b, k), np.float32)
    ids = np.empty((b, k), np.int64)
    rc = lib().oracle_scan_topk(_p(corpus), dtype, n, d, d, _p(q), b, k, score_mode, id_base,
                                _p(scores), _p(keys), _p(ids))
    if rc != 0:
        raise RuntimeError(f"oracle_scan_topk rc={rc}")
    return ids, scores, keys


def auto_merge(pairs, parent_of, child_count, prev_id, next_id, ratio_thresh=0.5, max_rounds=64):
    ids = np.array([p[0] for p in pairs], dtype=np.int64)
    sc = np.array([p[1] for p in pairs], dtype=np.float64)
    cap = max(16, 4 * len(pairs) + 16)
    out_ids = np.empty(cap, np.int64)
    out_sc = np.empty(cap, np.float64)
    arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (parent_of, child_count, prev_id, next_id)]
    n = lib().oracle_automerge(_p(ids), _p(sc), len(pairs), *[_p(a) for a in arrs], float(ratio_thresh),
                               int(max_rounds), _p(out_ids), _p(out_sc), cap)
    if n < 0:
        raise RuntimeError("oracle_automerge overflow")
    return [(int(out_ids[i]), float(out_sc[i])) for i in range(n)]


def fast_threads() -> int:
    return int(lib().oracle_fast_threads())


def use_all_host_threads() -> int:
    """Give the fast CPU arm every CPU this process may run on, whatever OMP_NUM_THREADS a launcher exported
    (torchrun sets it to 1 for its workers).  Returns the thread count now in use."""
    import os

    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        n = os.cpu_count() or 1
    lib().oracle_fast_set_threads(int(n))
    return fast_threads()


def fast_scan_topk(corpus_bits: np.ndarray, inv_norm: np.ndarray, queries: np.ndarray, k: int, id_base: int = 0):
    """Fast (timed) CPU arm: fp32 SIMD, all host threads, cosine only."""
    corpus_bits = np.ascontiguousarray(corpus_bits, dtype=np.uint16)
    inv_norm = np.ascontiguousarray(inv_norm, dtype=np.float32)
    q = np.ascontiguousarray(np.atleast_2d(queries), dtype=np.float32)
    n, d = corpus_bits.shape
    b = q.shape[0]
    scores = np.empty((b, k), np.float32)
    ids = np.empty((b, k), np.int64)
    rc = lib().oracle_fast_scan_topk(_p(corpus_bits), _p(inv_norm), n, d, _p(q), b, k, id_base, _p(scores), _p(ids))
    if rc != 0:
        raise RuntimeError(f"oracle_fast_scan_topk rc={rc}")
    return ids, scores
