"""ctypes binding of ``libtt_b200.so`` -- the only way the Python host code reaches the GPU kernels.

There is no fallback: if the library is missing or a call fails, this raises.
"""

from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libtt_b200.so")

OK = 0
ERR_TIMEOUT = -5
STATUS_RING_TIMEOUT, STATUS_EXCHANGE_TIMEOUT = 1, 2
DTYPE_BF16, DTYPE_F32 = 0, 1
SCORE_COSINE, SCORE_CHROMA_L2_EXP = 0, 1
SCAN_AUTO, SCAN_SIMT, SCAN_TCGEN05 = 0, 1, 2
MAX_SEGMENTS = 16

# every symbol include/tt_b200.h declares: name -> (restype, argtypes)
_P, _I, _L, _Z, _D = C.c_void_p, C.c_int, C.c_int64, C.c_size_t, C.c_double
SIGNATURES = {
    "tt_version": (_I, []),
    "tt_last_error": (C.c_char_p, []),
    "tt_stream_synchronize": (_I, [_P]),
    "tt_status_configure": (_I, [_P, _I]),
    "tt_status_read": (_I, [_P, _I]),
    "tt_scan_num_lists": (_I, [_I]),
    "tt_scan_max_kprime": (_I, []),
    "tt_prepare_queries": (_I, [_P, _I, _I, _P, _P, _P]),
    "tt_prepare_queries_rho": (_I, [_P, _I, _I, _P, _P, _P, _P]),
    "tt_certificate_credit": (_I, [_P, _I, _I, _P, C.c_float, _P]),
    "tt_scan_workspace_bytes": (_Z, []),
    "tt_scan_topk_bf16": (_I, [_P, _L, _I, _L, _P, _P, _P, _I, _I, _L, _I, _P, _P, _P, _P, _Z, _P]),
    "tt_scan_topk_bf16_segmented": (_I, [_P, _L, _I, _L, _P, _P, _P, _I, _I, _L, _P, _I, _P, _P, _P, _P, _Z, _P]),
    "tt_scan_gemm_workspace_bytes": (_Z, [_I, _I]),
    "tt_scan_gemm_topk_bf16": (_I, [_P, _L, _I, _L, _P, _P, _P, _I, _I, _L, _P, _P, _P, _P, _Z, _P]),
    "tt_rescore_workspace_bytes": (_Z, [_I, _I]),
    "tt_rescore_topk": (_I, [_P, _I, _L, _I, _L, _L, _P, _I, _P, _I, _P, _I, _I, _I, _P, _P, _P, _P, _P, _Z, _P]),
    "tt_scan_exact_workspace_bytes": (_Z, [_I, _I, _I]),
    "tt_scan_exact_f64": (_I, [_P, _I, _L, _I, _L, _L, _P, _I, _I, _I, _P, _P, _P, _P, _Z, _P]),
    "tt_scan_exact_f64_gated": (_I, [_P, _I, _L, _I, _L, _L, _P, _I, _I, _I, _P, _P, _P, _P, _P, _Z, _P]),
    "tt_merge_topk": (_I, [_P, _P, _I, _L, _L, _I, _I, _I, _I, _P, _P, _P]),
    "tt_rescore_topk_push": (_I, [_P, _I, _L, _I, _L, _L, _P, _I, _P, _I, _P, _I, _I, _I, _P, _P, _P, _P, _P, _Z, _P, _P, _P]),
    "tt_exchange_push": (_I, [_P, _Z, _P, _P]),
    "tt_peer_barrier": (_I, [_P, _P]),
    "tt_rescore_fused_workspace_bytes": (_Z, [_I, _I]),
    "tt_rescore_topk_fused": (_I, [_P, _I, _L, _I, _L, _L, _P, _I, _P, _I, _P, _I, _I, _I, _P, _P, _P, _P, _P, _Z, _P, _P, _P, _P,
                                   C.c_float, _P]),
    "tt_merge_topk_fused": (_I, [_P, _P, _I, _L, _L, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P]),
    "tt_merge_topk_pulled": (_I, [_P, _P, _I, _L, _L, _I, _I, _I, _I, _P, _P, _P, _P]),
    "tt_linear_bf16": (_I, [_P, _L, _I, _P, _I, _P, _P, _I, _P, _P]),
    "tt_layernorm_bf16": (_I, [_P, _L, _I, _P, _P, C.c_float, _P, _P]),
    "tt_embed_layernorm_bf16": (_I, [_P, _P, _L, _I, _P, _P, _P, _P, _P, C.c_float, _P, _P]),
    "tt_attention_varlen_bf16": (_I, [_P, _L, _I, _I, _P, _I, _I, _I, C.c_float, _P, _P]),
    "tt_cls_head_f32": (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _P, _P]),
    "tt_automerge_max_k": (_I, []),
    "tt_automerge": (_I, [_P, _P, _I, _I, _P, _P, _P, _P, _L, _D, _I, _P, _P, _P, _I, _P]),
}


MAX_PEERS = 16


class Exchange(C.Structure):
    """``tt_exchange_t`` (include/tt_b200.h): where one launch pushes its record and which flags it raises / waits on."""

    _fields_ = [("world", C.c_int), ("rank", C.c_int), ("epoch", C.c_uint32), ("rec_stride_bytes", C.c_uint64),
                ("ids_off_bytes", C.c_uint64), ("peer_recv", C.c_void_p * MAX_PEERS), ("peer_flags", C.c_void_p * MAX_PEERS),
                ("ticket", C.c_void_p), ("margins_off_bytes", C.c_uint64), ("epoch_dev", C.c_void_p),
                ("n_slots", C.c_uint32), ("slot_stride_bytes", C.c_uint64), ("flag_slot_stride", C.c_uint64)]


class AutomergeArgs(C.Structure):
    """``tt_automerge_args_t``: the tree arrays and outputs of stage 3, for the kernels that run it as their tail."""

    _fields_ = [("parent_of", C.c_void_p), ("child_count", C.c_void_p), ("prev_id", C.c_void_p), ("next_id", C.c_void_p),
                ("n_nodes", C.c_int64), ("ratio_thresh", C.c_double), ("max_rounds", C.c_int), ("out_ids", C.c_void_p),
                ("out_scores", C.c_void_p), ("out_len", C.c_void_p), ("max_out", C.c_int)]


class L2Cert(C.Structure):
    """``tt_l2_cert_t``: row-norm bounds that let the cosine-ordered shortlist certify a squared-L2 top-k."""

    _fields_ = [("row_norm_min", C.c_float), ("row_norm_max", C.c_float), ("eps", C.c_float)]


class TTError(RuntimeError):
    """A libtt_b200 entry point returned a negative code (message from ``tt_last_error``)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libtt_b200 error {code}: {message}")
        self.code = code


_lib = None
_lock = threading.RLock()  # re-entrant: status_word() configures the library under it


def lib():
    """The loaded library.  Raises if it has not been built -- there is no CPU path to fall back to."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python -m tensor_truth_b200.build` "
                        "(tensor_truth_b200 has no CPU fallback)")
                L = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(L, name)  # AttributeError if the .so does not export what the header declares
                    fn.restype = res
                    fn.argtypes = args
                _lib = L
    return _lib


# ---- device-side status (include/tt_b200.h, "Device-side status"): one pinned, device-mapped word per GPU that the
# kernels OR a TT_STATUS_* code into when a bounded wait runs out; polled after every host synchronisation.
_status_words: dict = {}
_status_np: dict = {}


def status_word(device_index: int):
    """The pinned status word of ``device_index`` (registered with the library on first use; that device must be current)."""
    w = _status_words.get(device_index)
    if w is None:
        import torch

        with _lock:
            w = _status_words.get(device_index)
            if w is None:
                w = torch.zeros(1, dtype=torch.int32).pin_memory()
                ms = int(os.environ.get("TT_WAIT_TIMEOUT_MS", "0"))
                with torch.cuda.device(device_index):
                    check(lib().tt_status_configure(w.data_ptr(), ms))
                _status_np[device_index] = w.numpy()  # polled after every host synchronisation: a plain memory read
                _status_words[device_index] = w
    return w


def set_wait_timeout_ms(device_index: int, ms: int) -> None:
    """Bound on one device-side ring wait (exchange waits: four times that); 0 restores the default (~4 s)."""
    import torch

    w = status_word(device_index)
    with torch.cuda.device(device_index):
        check(lib().tt_status_configure(w.data_ptr(), int(ms)))


def check_status(device_index: int) -> None:
    """Raise ``TTError(ERR_TIMEOUT)`` if a kernel on this GPU gave up a wait since the last check (call after the
    stream has been synchronised).  Clears the condition: the caller decides whether the index is still usable."""
    wn = _status_np.get(device_index)
    if wn is None or wn[0] == 0:
        return
    import torch

    w = _status_words[device_index]
    code = int(wn[0]) & 0xFFFFFFFF
    w.zero_()
    with torch.cuda.device(device_index):
        out = C.c_uint32(0)
        lib().tt_status_read(C.byref(out), 1)
    what = []
    if code & STATUS_RING_TIMEOUT:
        what.append("a TMA/MMA ring wait timed out inside a scan kernel")
    if code & STATUS_EXCHANGE_TIMEOUT:
        what.append(f"the peer exchange timed out waiting for rank {(code >> 8) & 0xFF}")
    raise TTError(ERR_TIMEOUT, "; ".join(what) or f"device status {code:#x}")


def check(rc: int) -> None:
    if rc != OK:
        raise TTError(rc, lib().tt_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device (or host) pointer of a torch tensor as an int (ctypes converts it for ``c_void_p`` parameters), or None."""
    return None if t is None else t.data_ptr()
