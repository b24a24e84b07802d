"""ctypes binding of libcmrag.so (the C ABI declared in include/cmrag.h).

This is the stub a maintainer of the reference would add.
There is NO CPU fallback: if the library is missing or no CUDA device is
present, calls raise.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

_PKG = Path(__file__).resolve().parent
# CMRAG_LIB: development override (kernel experiments built beside the product library)
LIB_PATH = Path(os.environ["CMRAG_LIB"]) if os.environ.get("CMRAG_LIB") else _PKG / "libcmrag.so"

CMR_OK, CMR_EINVAL, CMR_ECUDA, CMR_EWORKSPACE, CMR_EUNSUPPORTED = 0, -1, -2, -3, -4
CMR_FLAG_UNCERTIFIED = 1
CMR_MAX_K = 120
CMR_SLACK = 8
CMR_DENSE_AUTO, CMR_DENSE_SCAN, CMR_DENSE_MMA, CMR_DENSE_EXACT = 0, 1, 2, 3
CMR_BM25_AUTO, CMR_BM25_EXACT, CMR_BM25_HEAD, CMR_BM25_HEAD_NOFALLBACK = 0, 1, 2, 3

_lib = None

_vp, _i32, _i64, _f64, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_double, C.c_size_t

_SIGNATURES = {
    "cmr_last_error": (C.c_char_p, []),
    "cmr_version": (C.c_int, []),
    "cmr_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "cmr_dense_workspace_bytes": (_sz, [_i64, C.c_int, C.c_int, C.c_int]),
    "cmr_dense_topk": (C.c_int, [_vp, _i64, C.c_int, _vp, C.c_int, C.c_int, _vp, _i64, _f64,
                                 _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cmr_dense_topk_ex": (C.c_int, [_vp, _i64, C.c_int, _vp, C.c_int, C.c_int, _vp, _i64, _f64,
                                    _vp, _vp, _vp, _vp, _vp, _sz, _vp, C.c_int]),
    "cmr_f32_to_bf16": (C.c_int, [_vp, _vp, _i64, _vp]),
    "cmr_gather_rows": (C.c_int, [_vp, _i64, C.c_int, _i64, _vp, C.c_int, _vp, _vp]),
    "cmr_mmr_select": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _f64, _vp, _vp, _vp, _vp]),
    "cmr_hybrid_fuse": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, _f64, _f64,
                                  C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cmr_rrf_fuse": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp]),
    "cmr_filter_mask": (C.c_int, [_vp, _i64, C.c_int, _vp, _vp, C.c_int, _vp, _vp, _vp]),
    "cmr_masked_df": (C.c_int, [_vp, _vp, C.c_int, _i64, _vp, _vp, _vp, _vp]),
    "cmr_neardup_edges": (C.c_int, [_vp, _i64, C.c_int, C.c_float, C.c_int, C.c_int, _vp, C.c_uint64, _vp, _vp]),
    "cmr_neardup_rescore": (C.c_int, [_vp, C.c_int, _vp, C.c_uint64, _f64, _vp, _vp, _vp]),
    "cmr_neardup_resolve": (C.c_int, [_vp, C.c_uint64, _i64, _vp, _vp]),
    "cmr_jaccard_edges": (C.c_int, [_vp, _vp, C.c_int, _f64, _vp, C.c_uint64, _vp, _vp]),
    "cmr_shard_msg_bytes": (_sz, [C.c_int, C.c_int, C.c_int]),
    "cmr_shard_pack": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _i64, C.c_int, _i64,
                                 C.c_int, _vp, _vp]),
    "cmr_shard_merge": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp,
                                  _vp, _vp, _vp, _vp]),
    "cmr_tokenize_queries": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp]),
    "cmr_shard_exchange_pack": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _i64, C.c_int,
                                          _i64, C.c_int, _vp, _vp]),
    "cmr_shard_exchange_merge": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp,
                                           _vp, _vp, _vp, _vp, _vp, _vp]),
    "cmr_topk_merge": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "cmr_bm25_workspace_bytes": (_sz, [_vp, C.c_int, C.c_int]),
    "cmr_bm25_topk": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, _i64,
                                _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "cmr_bm25_topk_ex": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp, _i64,
                                   _vp, _vp, _vp, _vp, _vp, _sz, _vp, C.c_int]),
}


def load(build_if_missing: bool = True):
    """Load (building first if the .so is missing or stale and nvcc is present)."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_missing and not os.environ.get("CMRAG_LIB"):
        from . import build as _build
        try:
            if _build.needs_build():
                _build.build()
        except Exception as exc:  # no nvcc on this box: use the prebuilt library
            if not LIB_PATH.exists():
                raise RuntimeError(f"libcmrag.so missing and cannot be built: {exc}") from exc
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} not found: run `python -m classmate_rag_b200.build` "
                           "(the CUDA library is mandatory; there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def exported_symbols():
    return list(_SIGNATURES)


def last_error() -> str:
    return (load().cmr_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map a cmr_status to the exception type the reference would propagate."""
    if rc == CMR_OK:
        return
    msg = last_error()
    if rc == CMR_EINVAL:
        raise ValueError(msg)
    raise RuntimeError(f"libcmrag error {rc}: {msg}")
