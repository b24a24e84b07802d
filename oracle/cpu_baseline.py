"""CPU baseline for bench.py: the reference's hybrid path, ported faithfully.

TEST / MEASUREMENT INFRASTRUCTURE (see oracle/__init__.py).  This is what
``HybridRetriever.retrieve`` does per question on the CPU
(rag/retrieval/fusion.py:108-167):

  * dense: Chroma/hnswlib is not installable here, so the stand-in is an exact
    NumPy fp32 ``C @ q`` + partial sort (multi-threaded BLAS, all host cores) --
    labelled as a stand-in wherever it is reported;
  * MMR over the 24-row pool in NumPy fp32 (fusion.py:39-61);
  * BM25Store.search: filter scan, a BM25Okapi REBUILD over the candidate
    subset for every query, get_scores, full Python sort (bm25.py:175-212) --
    that per-query rebuild is the reference's behaviour and is kept;
  * rrf_fuse + per-id merge + final sort.
"""
from __future__ import annotations

import os
import time
from typing import List, Sequence

import numpy as np

from . import np_oracle as o


class ReferencePort:
    def __init__(self, emb_f32: np.ndarray, docs_tokens: Sequence[Sequence[str]]):
        self.emb = np.ascontiguousarray(emb_f32, dtype=np.float32)
        self.ids = [f"cm_{i:032x}" for i in range(self.emb.shape[0])]
        meta = {"language": "en"}
        self.entries = [(self.ids[i], list(t), meta) for i, t in enumerate(docs_tokens)]

    def _vector_search(self, q: np.ndarray, k: int, pool: int):
        sims = self.emb @ q.astype(np.float32)
        n = sims.shape[0]
        pool = min(pool, n)
        part = np.argpartition(-sims, pool - 1)[:pool]
        order = part[np.lexsort((part, -sims[part]))]
        cands = self.emb[order]
        # fusion.py:39-61 in NumPy fp32
        sims_q = (cands @ q.reshape(-1, 1).astype(np.float32)).ravel()
        sims_cc = cands @ cands.T
        sel = o.mmr_order(sims_q.astype(np.float64), sims_cc.astype(np.float64), k, 0.5)
        return [(self.ids[int(order[i])], float(1.0 - sims[order[i]])) for i in sel]

    def retrieve(self, q: np.ndarray, query_text: str, top_k: int = 10, k_vector: int = 8, k_bm25: int = 8):
        vec = self._vector_search(q, k_vector, max(k_vector, 24))
        bm = o.bm25_store_search(self.entries, query_text, None, k_bm25)
        return o.hybrid_merge(vec, bm, top_k)


def time_reference_port(port: ReferencePort, queries_f32: np.ndarray, query_texts: List[str], top_k: int,
                        budget_s: float = 25.0):
    """Run queries until the budget is spent; returns (seconds per query, n timed)."""
    times = []
    t_start = time.perf_counter()
    for q, text in zip(queries_f32, query_texts):
        t0 = time.perf_counter()
        port.retrieve(q, text, top_k)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s and len(times) >= 2:
            break
    return float(np.median(times)), len(times)


def host_cores() -> int:
    return os.cpu_count() or 1
