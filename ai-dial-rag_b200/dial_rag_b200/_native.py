"""ctypes binding of libdrag_b200.so (the C ABI declared in include/drag_b200.h).

The library is the ONLY compute path of this package: if it is missing or a call
fails, a ``DragError`` is raised -- there is no CPU / PyTorch fallback.
"""

from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libdrag_b200.so")

METRIC_CODES = {
    "cosine_sim": 0,
    "euclidean_dist": 1,
    "sqeuclidean_dist": 2,
    "inner_product": 3,
}
DTYPE_F32, DTYPE_BF16 = 0, 1
MAX_K = 2048
BATCH_MAX_K = 256      # drag_topk_batch limits (drag_topk_tc.cuh)
BATCH_MAX_DIM = 512


class DragError(RuntimeError):
    """A drag_b200 call returned a non-zero status."""

    def __init__(self, code: int, message: str):
        super().__init__(f"drag_b200 error {code}: {message}")
        self.code = code


class BertShape(C.Structure):
    _fields_ = [
        ("vocab", C.c_int32),
        ("hidden", C.c_int32),
        ("layers", C.c_int32),
        ("heads", C.c_int32),
        ("inter", C.c_int32),
        ("max_pos", C.c_int32),
        ("type_vocab", C.c_int32),
        ("ln_eps", C.c_float),
    ]


_P = C.c_void_p
_SIGNATURES = {
    "drag_last_error": (C.c_char_p, []),
    "drag_abi_version": (C.c_int, []),
    "drag_device_info": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "drag_encoder_create": (C.c_int, [C.POINTER(BertShape), C.POINTER(_P), C.c_int, C.c_int, C.c_int64, C.POINTER(_P)]),
    "drag_encoder_destroy": (C.c_int, [_P]),
    "drag_encoder_forward": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, _P]),
    "drag_encoder_embed_host": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "drag_encoder_forward_debug": (C.c_int, [_P, _P, _P, _P, C.c_int, C.c_int, _P, _P]),
    "drag_encoder_profile_begin": (C.c_int, [_P, C.c_int]),
    "drag_encoder_profile_end": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "drag_row_sqnorm": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int, _P, _P]),
    "drag_distances": (C.c_int, [C.c_int, _P, C.c_int, C.c_int64, C.c_int, _P, _P, C.c_int, _P, _P, _P]),
    "drag_topk_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "drag_topk": (
        C.c_int,
        [C.c_int, _P, C.c_int, C.c_int64, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int64,
         _P, _P, _P, _P, C.c_size_t, _P],
    ),
    "drag_topk_merge": (C.c_int, [C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "drag_rows_to_bf16": (C.c_int, [_P, C.c_int64, _P, _P]),
    "drag_row_norm_stats": (C.c_int, [_P, C.c_int64, _P, _P, _P]),
    "drag_topk_batch_workspace_bytes": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_size_t)]),
    "drag_topk_batch": (
        C.c_int,
        [C.c_int, _P, C.c_int, _P, C.c_int64, C.c_int, _P, _P, C.c_float, _P, C.c_int, C.c_int, C.c_int, C.c_int64,
         _P, _P, _P, _P, _P, C.c_size_t, _P],
    ),
    "drag_debug_tc_keys": (C.c_int, [C.c_int, _P, C.c_int64, C.c_int, _P, C.c_int, _P, C.c_int, _P, _P, C.c_size_t, _P]),
    "drag_rows_to_chunks": (C.c_int, [_P, C.c_int64, _P, C.c_int, _P, _P, _P, _P]),
    "drag_debug_gemm": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int, C.c_int,
                                  C.c_float, C.c_float, _P]),
    "drag_debug_mlp": (C.c_int, [C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int, C.c_float, _P]),
    "drag_debug_attention": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "drag_debug_attention_trace_words": (C.c_int, []),
    "drag_debug_set_attention_trace": (C.c_int, [C.c_int, _P]),
    "drag_wordpiece_create": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_P)]),
    "drag_wordpiece_destroy": (C.c_int, [_P]),
    "drag_wordpiece_encode": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
ABI_VERSION = 2   # include/drag_b200.h DRAG_ABI_VERSION (2: nine profile classes, drag_debug_mlp, attention trace entry points)

_lib = None
_lock = threading.Lock()


def load() -> C.CDLL:
    """Load the shared library once; raise DragError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise DragError(
                -1,
                f"{LIB_PATH} not found: build it with `python ai-dial-rag_b200/csrc/build.py` "
                "(no CPU fallback exists for this path)",
            )
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.drag_abi_version() != ABI_VERSION:   # a stale build: array sizes / signatures below would not match
            raise DragError(-1, f"{LIB_PATH} has ABI version {lib.drag_abi_version()}, this package needs {ABI_VERSION}: "
                                "rebuild it with `python ai-dial-rag_b200/csrc/build.py`")
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().drag_last_error()
        raise DragError(rc, msg.decode("utf-8", "replace") if msg else "unknown error")


def device_info(device: int = 0):
    lib = load()
    n, cc, sms = C.c_int(0), C.c_int(0), C.c_int(0)
    check(lib.drag_device_info(device, C.byref(n), C.byref(cc), C.byref(sms)))
    return {"n_devices": n.value, "compute_capability": cc.value, "sm_count": sms.value}
