"""ctypes loader for the C restatement (oracle/cmr_oracle.c).  TEST INFRASTRUCTURE."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB = _DIR / "_build" / "liboracle.so"
_lib = None


def build() -> Path:
    src = _DIR / "cmr_oracle.c"
    if not _LIB.exists() or _LIB.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_DIR)], check=True, capture_output=True)
    return _LIB


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB))
    return _lib


def exact_dots(q_bits: np.ndarray, c_bits: np.ndarray) -> np.ndarray:
    q = np.ascontiguousarray(q_bits, dtype=np.uint16)
    c = np.ascontiguousarray(c_bits, dtype=np.uint16)
    n, d = c.shape
    out = np.empty(n, dtype=np.float64)
    load().oracle_exact_dots(q.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p), C.c_int64(n), C.c_int(d),
                             out.ctypes.data_as(C.c_void_p))
    return out


def bm25_scores(term_ptr, post_doc, post_tf, doc_len, idf, avgdl, q_terms, k1=1.5, b=0.75) -> np.ndarray:
    term_ptr = np.ascontiguousarray(term_ptr, dtype=np.int64)
    post_doc = np.ascontiguousarray(post_doc, dtype=np.int32)
    post_tf = np.ascontiguousarray(post_tf, dtype=np.int32)
    doc_len = np.ascontiguousarray(doc_len, dtype=np.int32)
    idf = np.ascontiguousarray(idf, dtype=np.float64)
    q = np.ascontiguousarray(q_terms, dtype=np.int32)
    out = np.empty(doc_len.shape[0], dtype=np.float64)
    vp = C.c_void_p
    load().oracle_bm25_scores(term_ptr.ctypes.data_as(vp), post_doc.ctypes.data_as(vp), post_tf.ctypes.data_as(vp),
                              doc_len.ctypes.data_as(vp), idf.ctypes.data_as(vp), C.c_double(avgdl), C.c_double(k1),
                              C.c_double(b), q.ctypes.data_as(vp), C.c_int(len(q)), C.c_int32(idf.shape[0]),
                              C.c_int64(doc_len.shape[0]), out.ctypes.data_as(vp))
    return out
