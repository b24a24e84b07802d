"""Query tokenisation on the device (N3, ``cmr_tokenize_queries``): the batched form of
``tokenize(query, detect_lang_tag(query))`` + vocabulary lookup (reference
rag/retrieval/bm25.py:34-70,194-195).  The language tag is still decided on the host."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from .. import _lib
from .text import STOPWORDS_EN, STOPWORDS_IT


class TokenTableStruct(C.Structure):
    """Mirror of ``cmr_token_table`` (include/cmrag.h)."""
    _fields_ = [("fp", C.c_void_p), ("off", C.c_void_p), ("len", C.c_void_p), ("val", C.c_void_p),
                ("flags", C.c_void_p), ("pool", C.c_void_p), ("capacity", C.c_int32), ("reserved_", C.c_int32)]


_MASK64 = (1 << 64) - 1


def fnv1a64(data: bytes) -> int:
    h = 14695981039346656037
    for b in data:
        h = ((h ^ b) * 1099511628211) & _MASK64
    return h


class DeviceTokenizer:
    """Hash table of vocabulary + stopwords in device memory and the kernel call."""

    def __init__(self, vocab: Dict[str, int], device):
        self.device = torch.device(device)
        entries: Dict[bytes, list] = {}
        for word, tid in vocab.items():
            entries[word.encode("utf-8")] = [int(tid), 0]
        for words, bit in ((STOPWORDS_EN, 1), (STOPWORDS_IT, 2)):
            for w in words:
                e = entries.setdefault(w.encode("utf-8"), [-1, 0])
                e[1] |= bit
        cap = 8
        while cap < 2 * len(entries):
            cap *= 2
        fp = np.zeros(cap, dtype=np.uint64)
        off = np.zeros(cap, dtype=np.int32)
        ln = np.full(cap, -1, dtype=np.int32)
        val = np.full(cap, -1, dtype=np.int32)
        flags = np.zeros(cap, dtype=np.uint8)
        pool = bytearray()
        for key, (tid, fl) in entries.items():
            h = fnv1a64(key)
            slot = (h ^ (h >> 32)) & (cap - 1)
            while ln[slot] >= 0:
                slot = (slot + 1) & (cap - 1)
            fp[slot], off[slot], ln[slot], val[slot], flags[slot] = h, len(pool), len(key), tid, fl
            pool.extend(key)
        pool.extend(b"\0" * 16)
        dev = self.device
        self._arrays = [torch.from_numpy(fp.view(np.int64)).to(dev), torch.from_numpy(off).to(dev),
                        torch.from_numpy(ln).to(dev), torch.from_numpy(val).to(dev), torch.from_numpy(flags).to(dev),
                        torch.from_numpy(np.frombuffer(bytes(pool), dtype=np.uint8).copy()).to(dev)]
        self.struct = TokenTableStruct(*[t.data_ptr() for t in self._arrays], cap, 0)
        self.capacity = cap

    def __call__(self, queries: Sequence[str], *, langs: Optional[Sequence[str]] = None, max_terms: int = 32):
        """(q_terms int32 [B * max_terms] padded with -1, q_ptr int32 [B + 1], counts int32 [B])
        on the device.  ``langs``: per-query language tags (None = English for all)."""
        dev = self.device
        b = len(queries)
        raw = [q.encode("utf-8") for q in queries]
        ptr = np.zeros(b + 1, dtype=np.int64)
        ptr[1:] = np.cumsum([len(r) for r in raw])
        blob = np.frombuffer(b"".join(raw) + b"\0", dtype=np.uint8).copy()
        text = torch.from_numpy(blob).to(dev)
        text_ptr = torch.from_numpy(ptr).to(dev)
        lang_t = None
        if langs is not None:
            lang_t = torch.tensor([1 if (l or "").lower().startswith("it") else 0 for l in langs],
                                  dtype=torch.uint8, device=dev)
        out_terms = torch.empty((max(b, 1) * max_terms,), dtype=torch.int32, device=dev)
        out_counts = torch.zeros((max(b, 1),), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().cmr_tokenize_queries(
                text.data_ptr(), text_ptr.data_ptr(), None if lang_t is None else lang_t.data_ptr(), b,
                C.byref(self.struct), max_terms, out_terms.data_ptr(), out_counts.data_ptr(),
                torch.cuda.current_stream().cuda_stream))
        q_ptr = torch.arange(b + 1, dtype=torch.int32, device=dev) * max_terms
        return out_terms, q_ptr, out_counts[:b]
