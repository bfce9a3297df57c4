"""ctypes binding of ``libanr_b200.so`` (the C ABI declared in ``include/anr_b200.h``).

There is no CPU implementation behind this module: if the library is missing or no
sm_100 device is present the calls raise, they never fall back.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libanr_b200.so")

ANR_OK = 0
ERROR_NAMES = {1: "ANR_ERR_INVALID", 2: "ANR_ERR_CUDA", 3: "ANR_ERR_NO_DEVICE", 4: "ANR_ERR_OOM",
               5: "ANR_ERR_UNSUPPORTED"}


class AnrError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"{ERROR_NAMES.get(code, code)}: {message}")
        self.code = code


_P = C.c_void_p
_I32 = C.c_int32
_I64 = C.c_int64
_F64 = C.c_double

# name -> argtypes (every function returns int except the two noted below)
SIGNATURES = {
    "anr_ctx_create": [_I32, C.POINTER(_P)],
    "anr_ctx_destroy": [_P],
    "anr_ctx_sync": [_P],
    "anr_ctx_set_beside_dense": [_P, _I32],
    "anr_ctx_info": [_P, C.POINTER(_I32), C.POINTER(_I64), C.POINTER(_I64)],
    "anr_ctx_last_rerun": [_P, C.POINTER(_I32), C.POINTER(_I32)],
    "anr_set_option": [C.c_char_p, _I32],
    "anr_ctx_profile_enable": [_P, _I32],
    "anr_ctx_profile_read": [_P, _I32, C.POINTER(_F64), C.POINTER(_I64)],
    "anr_ctx_timeline_enable": [_P, _I32],
    "anr_ctx_timeline_read": [_P, C.POINTER(_F64)],
    "anr_dense_create": [_P, _P, _I64, _I32, _I32, C.POINTER(_P)],
    "anr_dense_upload": [_P, _P, _I64, _P, _I64],
    "anr_dense_destroy": [_P],
    "anr_dense_shape": [_P, C.POINTER(_I64), C.POINTER(_I32)],
    "anr_dense_set_shadow": [_P, _I32],
    "anr_dense_invalidate": [_P],
    "anr_dense_search": [_P, _P, _P, _I32, _I32, _P, _I64, _P, _P, _P, _P],
    "anr_bm25_create": [_P, _P, _P, _P, _P, _P, _I32, _I32, _F64, _F64, _F64, C.POINTER(_P)],
    "anr_bm25_reweight": [_P, _P, _P, _P, _P, _F64, _F64, _F64],
    "anr_bm25_destroy": [_P],
    "anr_bm25_shape": [_P, C.POINTER(_I32), C.POINTER(_I32), C.POINTER(_I64)],
    "anr_bm25_search": [_P, _P, _P, _P, _I32, _I32, _P, _P, _I64, _P, _P, _P, _P],
    "anr_bm25_scores": [_P, _P, _P, _I32, _P, _P],
    "anr_wrrf_fuse": [_P, _P, _P, _P, _I32, _I32, _I32, _F64, _I32, _P, _P, _P, _P],
    "anr_hybrid_search": [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _P, _P, _P, _I64, _F64, _F64,
                          _F64, _I32, _P, _P, _P, _P, _P, _P, _P, _P],
    "anr_dense_search_keys": [_P, _P, _P, _I32, _I32, _P, _I64, _P, _P],
    "anr_bm25_search_keys": [_P, _P, _P, _P, _I32, _I32, _P, _P, _I64, _P, _P],
    "anr_hybrid_search_keys": [_P, _P, _P, _P, _P, _P, _I32, _I32, _P, _P, _I64, _I64, _P, _P],
    "anr_topk_merge": [_P, _P, _I32, _I32, _I32, _P, _P, _P, _P],
    "anr_sharded_fuse": [_P, _P, _I32, _I32, _I32, _F64, _F64, _F64, _I32, _P, _P, _P, _P],
    "anr_sqlite_read_blobs": [C.c_char_p, C.c_char_p, _P, _I64, _I64, _P, C.POINTER(_I64),
                              C.POINTER(_I32)],
}
EXPORTS = sorted(list(SIGNATURES) + ["anr_abi_version", "anr_last_error"])

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Loads the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python a-nice-rag_b200/build.py` "
            "(there is no CPU fallback for the retrieval kernels)")
    lib = C.CDLL(LIB_PATH)
    lib.anr_abi_version.restype = C.c_int
    lib.anr_abi_version.argtypes = []
    lib.anr_last_error.restype = C.c_char_p
    lib.anr_last_error.argtypes = []
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = argtypes
    if lib.anr_abi_version() != 1:
        raise ImportError(f"{LIB_PATH}: ABI version {lib.anr_abi_version()} != 1")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != ANR_OK:
        raise AnrError(rc, load().anr_last_error().decode("utf-8", "replace"))


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args))


def ptr(x) -> Optional[int]:
    """Address of a numpy array / torch tensor / raw int; None stays NULL."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):       # torch tensor (host or device)
        return x.data_ptr()
    return x.ctypes.data             # numpy array
