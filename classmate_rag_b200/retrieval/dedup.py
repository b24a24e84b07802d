"""Drop-in for ``rag.utils.dedup`` (reference rag/utils/dedup.py:40-55, called by ingest_file,
rag/pipeline/rag.py:308-324): near-duplicate chunk filtering by Jaccard similarity of token
5-gram shingle sets, greedy keep-first.

The host does the string work once per chunk (normalise, shingle, map every distinct shingle
tuple to an integer through a dictionary -- equal id <=> equal shingle, so nothing is hashed
away).  The quadratic part -- set intersections of every pair and the threshold test -- runs
on the GPU (``cmr_jaccard_edges``), and the greedy pass over the surviving edges is the same
``cmr_neardup_resolve`` the embedding near-duplicate filter uses.  The decision is integer
counts and one float64 division per pair: identical to the reference (golden vectors from the
live reference: tests/golden/reference_dedup.json).  No CPU fallback.
"""
from __future__ import annotations

import re
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

from .. import _lib, neardup

_PUNCT = re.compile(r"[^\w\s]", re.UNICODE)
SHINGLE = 5


def norm_tokens(text: str) -> List[str]:
    """Lower-case, every non-word non-space character becomes a space, split on whitespace."""
    return _PUNCT.sub(" ", (text or "").lower()).split()


def shingle_sets(blocks: Sequence[str], k: int = SHINGLE) -> Tuple[np.ndarray, np.ndarray]:
    """CSR of the blocks' shingle sets: (ptr int32 [n+1], items int32), ids ascending and
    unique inside a set.  A text shorter than k tokens is one shingle; an empty text has none."""
    ids: Dict[Tuple[str, ...], int] = {}
    ptr = np.zeros(len(blocks) + 1, dtype=np.int32)
    parts: List[np.ndarray] = []
    for i, text in enumerate(blocks):
        toks = tuple(norm_tokens(text))
        if not toks:
            grams: List[Tuple[str, ...]] = []
        elif len(toks) < k:
            grams = [toks]
        else:
            grams = [toks[j:j + k] for j in range(len(toks) - k + 1)]
        mine = np.unique(np.fromiter((ids.setdefault(g, len(ids)) for g in grams), dtype=np.int32, count=len(grams)))
        parts.append(mine)
        ptr[i + 1] = ptr[i] + mine.size
    items = np.concatenate(parts) if parts else np.zeros(0, dtype=np.int32)
    return ptr, items.astype(np.int32, copy=False)


def jaccard_edges(ptr: torch.Tensor, items: torch.Tensor, threshold: float) -> torch.Tensor:
    """Edges (i << 32 | j, j < i, Jaccard >= threshold) as an int64 device tensor, unordered."""
    if not ptr.is_cuda or ptr.dtype != torch.int32 or items.dtype != torch.int32:
        raise RuntimeError("shingle sets must be CUDA int32 tensors (no CPU path)")
    n = ptr.numel() - 1
    dev = ptr.device
    lib = _lib.load()
    count = torch.zeros((1,), dtype=torch.int64, device=dev)
    cap = max(1024, 4 * n)
    with torch.cuda.device(dev):
        while True:
            edges = torch.empty((cap,), dtype=torch.int64, device=dev)
            _lib.check(lib.cmr_jaccard_edges(ptr.data_ptr(), items.data_ptr() if items.numel() else None, n,
                                             float(threshold), edges.data_ptr(), cap, count.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream))
            found = int(count.item())
            if found <= cap:
                return edges[:found]
            cap = found + 1024   # the buffer overflowed: the exact size is now known


def dedup_keep_mask(blocks: Sequence[str], *, jaccard_threshold: float = 0.92, device=None) -> np.ndarray:
    """bool [len(blocks)]: True where dedup_text_blocks keeps the block."""
    if not torch.cuda.is_available():
        raise RuntimeError("dedup needs a CUDA device: the product has no CPU path")
    n = len(blocks)
    if n == 0:
        return np.zeros(0, dtype=bool)
    dev = torch.device("cuda" if device is None else device)
    ptr, items = shingle_sets(blocks)
    edges = jaccard_edges(torch.from_numpy(ptr).to(dev), torch.from_numpy(items).to(dev), jaccard_threshold)
    return neardup.resolve(edges, n).cpu().numpy().astype(bool)


def dedup_text_blocks(blocks: List[str], *, jaccard_threshold: float = 0.92) -> List[str]:
    """Preserve order; drop any block whose shingle Jaccard with a previously kept block is
    >= jaccard_threshold (same signature and result as the reference's function)."""
    keep = dedup_keep_mask(blocks, jaccard_threshold=jaccard_threshold)
    return [b for b, k in zip(blocks, keep) if k]
